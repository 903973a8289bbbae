from oracle.dwm_numpy import DWMFlowSimulation  # noqa: F401
