class XYView:
    def __init__(self, *a, **k):
        pass


EastNorthView = XYView
