"""-C / -N plugin modules with the reference's contracts (REF/cGP/cluster.py:3-15, REF/cGP/classify.py:3-9)."""
