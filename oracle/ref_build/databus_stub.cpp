// Single-process stand-in for gcm/system/DataBus.cpp (oracle/_ref build only -- test
// infrastructure).  The real DataBus needs MPI-2 C++ bindings, absent from this image.
// Everything here is the one-rank behaviour of the reference:
//   - sync_nodes          = the same-rank branch of DataBus::sync_nodes (DataBus.cpp:64-79)
//   - get_max_possible_tau = identity (Allreduce(MIN) over one rank, DataBus.cpp:48-57)
//   - the collective-only members are no-ops.
#include "system/DataBus.h"

DataBus::DataBus()
{
	data_bus_type.assign("Single-process DataBus stub");
	proc_num = 0;
	procs_total_num = 1;
	logger->set_proc_num(proc_num);
	mesh_set = TetrMeshSet::getInstance();
	collision_detector = NULL;
	local_numbers = NULL;
	MPI_NODE_TYPES = NULL;
}

DataBus::~DataBus() { }

string* DataBus::get_data_bus_type() { return &data_bus_type; }

void DataBus::attach(CollisionDetector* cd) { collision_detector = cd; }

float DataBus::get_max_possible_tau(float local_time_step) { return local_time_step; }

int DataBus::sync_nodes()
{
	for (int i = 0; i < mesh_set->get_number_of_local_meshes(); i++)
	{
		TetrMesh_1stOrder* mesh = mesh_set->get_local_mesh(i);
		for (int j = 0; j < (int)mesh->nodes.size(); j++)
			if (mesh->nodes[j].isRemote ())
			{
				ElasticNode *node = mesh_set->get_mesh_by_zone_num(mesh->nodes[j].remote_zone_num)->get_node(mesh->nodes[j].remote_num);
				memcpy(mesh->nodes[j].values, node->values, 13*sizeof(float));
				memcpy(mesh->nodes[j].coords, node->coords, 3*sizeof(float));
			}
	}
	return 0;
}

void DataBus::load_zones_info(vector<int>* map)
{
	zones_info.clear();
	for(int i = 0; i < (int)map->size(); i++)
		zones_info.push_back( map->at(i) );
}

int DataBus::get_proc_num() { return proc_num; }
int DataBus::get_procs_total_num() { return procs_total_num; }
int DataBus::get_proc_for_zone(int zone_num) { return zones_info[zone_num]; }
void DataBus::sync_outlines() { }
void DataBus::terminate() { }
void DataBus::create_custom_types() { }
void DataBus::sync_faces_in_intersection(MeshOutline **intersections, int **fs, int **fl) { }
void DataBus::sync_tetrs() { }

DataBus* DataBus::getInstance()
{
	static DataBus db;
	return &db;
}
