from oracle.dwm_numpy import MannTurbulenceField, RandomTurbulence  # noqa: F401
