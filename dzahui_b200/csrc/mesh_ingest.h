// mesh_ingest.h — host half of MeshBuilder::build_mesh_1d (reference mesh/mesh_builder.rs:58-84,140-190,204-277):
// text -> sorted node coordinates + the scalars the vertex generator needs.  The GL vertex / index / element
// arrays themselves are generated on the device from the sorted nodes (mesh_kernels in api.cu).
#pragma once
#include <string>
#include <vector>

namespace dzb {

enum class IngestStatus { Ok, Io, Parse };

struct Ingested1D {
  std::vector<double> nodes;  // the kept coordinate of every `v` line, ascending (stable)
  int kept_axis = 0;          // 0 = x, 1 = y, 2 = z
  double max_length = 0.0;    // -first + last                  (mesh_builder.rs:270)
  double prom_width = 0.0;    // max_length*6/(6n-6)*multiplier (mesh_builder.rs:277)
  float translation[3] = {0.f, 0.f, 0.f};  // model matrix translation (mesh_builder.rs:310-318)
  std::string message;
};

IngestStatus ingest_obj_1d(const char* path, bool has_multiplier, double multiplier, Ingested1D* out);

}  // namespace dzb
