def _noop(*a, **k):
    return None


ion = close = gca = pcolormesh = gcf = _noop


def subplots(*a, **k):
    return None, None
