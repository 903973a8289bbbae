class HAWC2WindTurbines:
    def __init__(self, *a, **k):
        raise NotImplementedError("HAWC2 co-simulation is out of scope (SURVEY.md section 2 #17)")
