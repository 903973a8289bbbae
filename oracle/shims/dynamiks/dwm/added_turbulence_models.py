class SynchronizedAutoScalingIsotropicMannTurbulence:
    pass


class AutoScalingIsotropicMannTurbulence:
    pass
