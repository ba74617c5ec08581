// empty VTK stub (oracle/_ref build only); see vtkUnstructuredGrid.h
