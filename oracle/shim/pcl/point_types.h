// oracle shim (test infrastructure only): pcl::PointXYZI with PCL's 32-byte layout
// (xyz + pad float, intensity + 3 pad floats). The harness stashes the input index in data[3].
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
namespace pcl {
struct alignas(16) PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  union { struct { float intensity; }; float data_c[4]; };
  PointXYZI() { data[0] = data[1] = data[2] = 0.f; data[3] = 1.f; data_c[0] = data_c[1] = data_c[2] = data_c[3] = 0.f; }
};
static_assert(sizeof(PointXYZI) == 32, "PointXYZI must be 32 B like PCL's");
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
struct PCLPointCloud2 {};
}
