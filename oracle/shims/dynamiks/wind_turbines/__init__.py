from oracle.dwm_numpy import PyWakeWindTurbines  # noqa: F401
