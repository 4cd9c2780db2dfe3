/* render.cu — placeholder until the wavefront integrator lands (next commit). */
#include "device.cuh"
using namespace gpurt;
struct gpurt_pipe { int dummy; };
extern "C" {
#define NOT_YET return set_error("integrator not built yet"), GPURT_E_STATE
int gpurt_pipe_create(gpurt_scene*, gpurt_accel*, gpurt_pipe**) { NOT_YET; }
int gpurt_pipe_destroy(gpurt_pipe*) { return GPURT_OK; }
int gpurt_pipe_reset_frame(gpurt_pipe*) { NOT_YET; }
int gpurt_pipe_render_frame(gpurt_pipe*, const GpurtPipeParams*, const GpurtCamera*, uint32_t, uint32_t) { NOT_YET; }
int gpurt_pipe_frame_index(const gpurt_pipe*, int32_t*) { NOT_YET; }
int gpurt_pipe_read_image(gpurt_pipe*, float*, int) { NOT_YET; }
int gpurt_pipe_read_gbuffer(gpurt_pipe*, int, float*, int) { NOT_YET; }
int gpurt_pipe_ray_counts(const gpurt_pipe*, uint64_t*) { NOT_YET; }
int gpurt_pipe_device_image(gpurt_pipe*, void**) { NOT_YET; }
int gpurt_tonemap(gpurt_pipe*, int, float, float, uint8_t*, int) { NOT_YET; }
}
