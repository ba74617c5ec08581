// Include-path stub for the VTK headers pulled in by TetrMesh_1stOrder.h
// (oracle/_ref build only). The only code that touches these types is the
// .vtu loader, which the oracle driver never calls; every method aborts.
#ifndef GCM_ORACLE_STUB_VTK_H
#define GCM_ORACLE_STUB_VTK_H
#include <cstdlib>
struct vtkDataArrayStub {
	void SetNumberOfComponents(int) { abort(); }
	double GetValue(int) { abort(); return 0; }
	void GetTupleValue(int, double*) { abort(); }
};
typedef vtkDataArrayStub vtkDoubleArray;
typedef vtkDataArrayStub vtkIntArray;
struct vtkPointData { vtkDataArrayStub* GetArray(const char*) { abort(); return 0; } };
struct vtkTetra { int GetPointId(int) { abort(); return 0; } };
struct vtkUnstructuredGrid {
	static vtkUnstructuredGrid* New() { abort(); return 0; }
	int GetNumberOfPoints() { abort(); return 0; }
	int GetNumberOfCells() { abort(); return 0; }
	vtkPointData* GetPointData() { abort(); return 0; }
	double* GetPoint(int) { abort(); return 0; }
	void* GetCell(int) { abort(); return 0; }
};
struct vtkXMLUnstructuredGridReader {
	static vtkXMLUnstructuredGridReader* New() { abort(); return 0; }
	void SetFileName(char*) { abort(); }
	void Update() { abort(); }
	vtkUnstructuredGrid* GetOutput() { abort(); return 0; }
};
#endif
