from oracle.dwm_numpy import TurbulenceFieldSite  # noqa: F401
