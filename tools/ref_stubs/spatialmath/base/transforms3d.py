def isrot(*a, **k):
    raise NotImplementedError("spatialmath stub")
