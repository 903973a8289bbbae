"""Stand-in package tree: every name re-exports the restatement in oracle/dwm_numpy.py."""
