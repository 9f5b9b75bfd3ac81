// temporary stubs
#include "stages.h"
namespace bshot {
#define NI(name) { set_error(name ": not implemented yet"); return BSHOT_E_STATE; }
int grid_build(Ctx*, size_t, int) NI("grid_build")
int detect_seg_ratio(Ctx*, float, int, int) NI("detect_seg_ratio")
int detect_topk(Ctx*, int) NI("detect_topk")
int normals_query(Ctx*, const float4*, size_t, float, int, float4*) NI("normals_query")
int normals_compute(Ctx*, int, float, int) NI("normals_compute")
int shot_compute(Ctx*, float, bool, bool) NI("shot_compute")
int binarize(Ctx*, const float*, size_t, uint64_t*) NI("binarize")
int frame_run(Ctx*, const bshot_params*) NI("frame_run")
}
