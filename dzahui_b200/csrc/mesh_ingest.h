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

// Host half of MeshBuilder::build_mesh_2d (mesh/mesh_builder.rs:342-466, obj_face_checker :89-129): text -> the two kept
// coordinates of every vertex, the 0-based triangle list, and the viewing scalars.  Boundary detection runs on the device.
struct Ingested2D {
  std::vector<double> xy;           // 2 per vertex, file order
  std::vector<unsigned> triangles;  // 3 per triangle
  int constant_axis = 2;
  double max_length = 0.0;                 // :455-461
  float translation[3] = {0.f, 0.f, 0.f};  // :464-465, :483-487
  std::string message;
};
IngestStatus ingest_obj_2d(const char* path, Ingested2D* out);

}  // namespace dzb
