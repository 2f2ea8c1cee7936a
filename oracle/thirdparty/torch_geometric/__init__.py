"""Restated subset of torch_geometric==2.0.3 (requirements.txt:8). Oracle only."""
__version__ = "2.0.3+oracle-restatement"
