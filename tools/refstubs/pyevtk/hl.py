def gridToVTK(*args, **kwargs):
    raise NotImplementedError("pyevtk is not installed in this environment")


pointsToVTK = gridToVTK
