// Include-path stub for <mpi.h>: lets the reference's headers parse without MPI.
// Test infrastructure only (oracle/_ref build). No MPI call is ever made: the
// translation units that use MPI (DataBus.cpp & co) are replaced by databus_stub.cpp.
#ifndef GCM_ORACLE_STUB_MPI_H
#define GCM_ORACLE_STUB_MPI_H
namespace MPI {
	class Comm {};
	class Errhandler {};
	class Datatype {};
	class Request {};
	class Status {};
	typedef long Aint;
}
#endif
