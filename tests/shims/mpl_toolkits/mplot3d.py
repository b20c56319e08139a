Axes3D = object
