/* spatial_sim_c.h — flat C binding of the SpatialSimulator class
 * (spatial_sbml_b200/csrc/host/spatialsimulator.h), for callers that cannot link C++: the role the
 * reference's SWIG layer plays (reference bindings/swig/spatialSBML.i wraps
 * SpatialSBML/spatialsimulator.h:49-157).  One function per public method, same argument meaning.
 *
 * Borrowed pointers (ssb_get_variable, ssb_get_geometry, ...) stay valid until the simulator is
 * re-initialised or freed.  Functions returning int give 0 on success, negative on failure with the
 * message in ssb_last_error().  There is no CPU fallback: without a CUDA device ssb_init_* fails
 * (unless ssb_set_host_only was called, in which case only the set-up can be inspected).
 */
#ifndef SPATIAL_SIM_C_H_
#define SPATIAL_SIM_C_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SSB_API __attribute__((visibility("default")))
#else
#define SSB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssb_sim ssb_sim;

SSB_API ssb_sim* ssb_new(void);                                   /* SpatialSimulator(), spatialsimulator.h:49 */
SSB_API void ssb_free(ssb_sim* s);
SSB_API const char* ssb_last_error(ssb_sim* s);
SSB_API int ssb_is_initialized(ssb_sim* s);

SSB_API void ssb_set_device(ssb_sim* s, int ordinal);             /* before init */
SSB_API void ssb_set_host_only(ssb_sim* s, int host_only);        /* before init */
SSB_API void ssb_set_slab(ssb_sim* s, int lo, int hi);            /* before init: owned rows of the slab axis */
/* before init: several GPUs behind one simulator - one slab of the grid per entry (an ordinal may repeat), stepped by the
 * library's sharded step with peer copies of the halo rows; results are bit-identical to one device */
SSB_API void ssb_set_devices(ssb_sim* s, const int* ordinals, int n);
SSB_API int ssb_get_num_slabs(ssb_sim* s);
SSB_API void ssb_get_slab_rows(ssb_sim* s, int* held_lo, int* own_lo, int* own_hi);  /* rows of the whole grid; zeros when not a slab */
SSB_API void ssb_set_cuda_stream(ssb_sim* s, void* cuda_stream);

SSB_API int ssb_init_from_file(ssb_sim* s, const char* file, int xdim, int ydim, int zdim);   /* :66 */
SSB_API int ssb_init_from_string(ssb_sim* s, const char* sbml, int xdim, int ydim, int zdim); /* :78 */

/* :88 — returns initialTime + step (NaN on failure) */
SSB_API double ssb_one_step(ssb_sim* s, double initial_time, double step);
SSB_API int ssb_run_steps(ssb_sim* s, double first_step_index, double step, int nsteps);
/* oneStep with the values of n species taken from / returned to caller-owned compact host arrays (nx*ny*nz doubles each,
 * ideally pinned; in[i] / out[i] may be NULL); upload, compute and download are pipelined over nslabs slabs (<= 0: 16) */
SSB_API double ssb_one_step_host(ssb_sim* s, double initial_time, double step, int n, const char* const* ids,
                                 const double* const* in, double* const* out, int nslabs);
SSB_API int ssb_run(ssb_sim* s, double end_time, double step);                                /* :286 */

SSB_API void ssb_set_parameter(ssb_sim* s, const char* id, double value);                     /* :107 */
SSB_API void ssb_set_parameter_uniformly(ssb_sim* s, const char* id, double value);           /* :97 */
SSB_API int ssb_get_index_for_position(ssb_sim* s, double x, double y);                       /* :116 */
SSB_API int ssb_get_variable_length(ssb_sim* s);                                              /* :123 */
SSB_API int ssb_get_geometry_length(ssb_sim* s);                                              /* :130 */
SSB_API int ssb_get_xdim(ssb_sim* s);
SSB_API int ssb_get_ydim(ssb_sim* s);
SSB_API int ssb_get_zdim(ssb_sim* s);
SSB_API double ssb_get_variable_at(ssb_sim* s, const char* id, int x, int y, int z);          /* :157 */
SSB_API void ssb_flip_volume_order(ssb_sim* s);                                               /* :187 */

SSB_API double* ssb_get_variable(ssb_sim* s, const char* id, int* length);                    /* :209 */
/* the array behind ssb_get_variable is kept current after every step (and written values are picked up before the next one),
 * as the reference's pointer into its state is; ssb_release_variable ends that for `id` (the pointer stays valid) */
SSB_API void ssb_release_variable(ssb_sim* s, const char* id);
SSB_API int* ssb_get_geometry(ssb_sim* s, const char* compartment, int* length);              /* :220 */
SSB_API uint8_t* ssb_get_boundary_type(ssb_sim* s, const char* compartment, int* length);     /* :231, 6 bools per point */
SSB_API int* ssb_get_boundary(ssb_sim* s, const char* compartment, int* length);              /* :242 */
SSB_API double* ssb_get_x(ssb_sim* s, int* length);                                           /* :251 */
SSB_API double* ssb_get_y(ssb_sim* s, int* length);
SSB_API double* ssb_get_z(ssb_sim* s, int* length);

/* additions */
SSB_API int ssb_get_variable_compact(ssb_sim* s, const char* id, double* out, size_t n);
SSB_API int ssb_set_variable_compact(ssb_sim* s, const char* id, const double* in, size_t n);
SSB_API double ssb_get_stable_time_step(ssb_sim* s, double dt);   /* checkStability.cpp reduction */
SSB_API long ssb_get_num_domain_cells(ssb_sim* s);
SSB_API int ssb_get_num_staged_species(ssb_sim* s);
SSB_API int ssb_get_time_slot(ssb_sim* s);                      /* constants-table slot of the `time` symbol, -1: the model does not use it */
SSB_API long ssb_get_device_launch_count(ssb_sim* s);
SSB_API void* ssb_get_device_handle(ssb_sim* s);                  /* sb2_sim* of include/spatial_b200.h */
/* after stepping the device handle directly (split-phase API of the slab-sharded path): host mirrors are out of date */
SSB_API void ssb_device_state_changed(ssb_sim* s);
/* text is copied into buf (NUL terminated, truncated to buflen); returns the full length */
SSB_API int ssb_get_device_program_text(ssb_sim* s, char* buf, int buflen);
/* model-specialised stage kernel (spatial_b200.h, sb2_specialized / sb2_specialize_info) */
SSB_API int ssb_is_specialized(ssb_sim* s);
SSB_API int ssb_get_specialize_info(ssb_sim* s, char* buf, int buflen);
SSB_API int ssb_specialize_now(ssb_sim* s);   /* 1: specialised kernel in place, 0: interpreter kernel stays */
SSB_API int ssb_get_setup_text(ssb_sim* s, const char* what, char* buf, int buflen);

#ifdef __cplusplus
}
#endif
#endif /* SPATIAL_SIM_C_H_ */
