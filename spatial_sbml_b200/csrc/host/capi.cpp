// capi.cpp — flat C binding of SpatialSimulator (include/spatial_sim_c.h).
#include "../../../include/spatial_sim_c.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include "spatialsimulator.h"

struct ssb_sim {
  SpatialSimulator sim;
  std::string error;
};

namespace {
int copyText(const std::string& t, char* buf, int buflen) {
  if (buf && buflen > 0) snprintf(buf, (size_t)buflen, "%s", t.c_str());
  return (int)t.size();
}
int initResult(ssb_sim* s) {
  if (s->sim.isInitialized()) { s->error.clear(); return 0; }
  s->error = s->sim.getLastError().empty() ? "initialisation failed" : s->sim.getLastError();
  return -1;
}
}  // namespace

extern "C" {

ssb_sim* ssb_new(void) { return new ssb_sim; }
void ssb_free(ssb_sim* s) { delete s; }
const char* ssb_last_error(ssb_sim* s) { return s ? s->error.c_str() : "null simulator"; }
int ssb_is_initialized(ssb_sim* s) { return s && s->sim.isInitialized() ? 1 : 0; }
void ssb_set_device(ssb_sim* s, int ordinal) { if (s) s->sim.setDevice(ordinal); }
void ssb_set_host_only(ssb_sim* s, int h) { if (s) s->sim.setHostOnly(h != 0); }
void ssb_set_slab(ssb_sim* s, int lo, int hi) { if (s) s->sim.setSlab(lo, hi); }
void ssb_set_devices(ssb_sim* s, const int* ordinals, int n) { if (s) s->sim.setDevices(ordinals, n); }
int ssb_get_num_slabs(ssb_sim* s) { return s ? s->sim.getNumSlabs() : 0; }
void ssb_get_slab_rows(ssb_sim* s, int* held_lo, int* own_lo, int* own_hi) {
  int a = 0, b = 0, c = 0;
  if (s) s->sim.getSlabRows(a, b, c);
  if (held_lo) *held_lo = a;
  if (own_lo) *own_lo = b;
  if (own_hi) *own_hi = c;
}
void ssb_set_cuda_stream(ssb_sim* s, void* st) { if (s) s->sim.setCudaStream(st); }

int ssb_init_from_file(ssb_sim* s, const char* file, int x, int y, int z) {
  if (!s || !file) return -1;
  try { s->sim.initFromFile(file, x, y, z); } catch (const std::exception& e) { s->error = e.what(); return -1; }
  return initResult(s);
}
int ssb_init_from_string(ssb_sim* s, const char* sbml, int x, int y, int z) {
  if (!s || !sbml) return -1;
  try { s->sim.initFromString(sbml, x, y, z); } catch (const std::exception& e) { s->error = e.what(); return -1; }
  return initResult(s);
}

double ssb_one_step(ssb_sim* s, double t, double dt) {
  if (!s) return NAN;
  try { return s->sim.oneStep(t, dt); } catch (const std::exception& e) { s->error = e.what(); return NAN; }
}
double ssb_one_step_host(ssb_sim* s, double t, double dt, int n, const char* const* ids, const double* const* in, double* const* out, int nslabs) {
  if (!s) return NAN;
  try { return s->sim.oneStepHost(t, dt, n, ids, in, out, nslabs); } catch (const std::exception& e) { s->error = e.what(); return NAN; }
}
int ssb_run_steps(ssb_sim* s, double first, double dt, int n) {
  if (!s) return -1;
  try { s->sim.runSteps(first, dt, n); } catch (const std::exception& e) { s->error = e.what(); return -1; }
  return 0;
}
int ssb_run(ssb_sim* s, double end, double dt) {
  if (!s) return -1;
  try { s->sim.run(end, dt); } catch (const std::exception& e) { s->error = e.what(); return -1; }
  return 0;
}

void ssb_set_parameter(ssb_sim* s, const char* id, double v) { if (s && id) s->sim.setParameter(id, v); }
void ssb_set_parameter_uniformly(ssb_sim* s, const char* id, double v) { if (s && id) s->sim.setParameterUniformly(id, v); }
int ssb_get_index_for_position(ssb_sim* s, double x, double y) { return s ? s->sim.getIndexForPosition(x, y) : -1; }
int ssb_get_variable_length(ssb_sim* s) { return s ? s->sim.getVariableLength() : 0; }
int ssb_get_geometry_length(ssb_sim* s) { return s ? s->sim.getGeometryLength() : 0; }
int ssb_get_xdim(ssb_sim* s) { return s ? s->sim.getXDim() : 0; }
int ssb_get_ydim(ssb_sim* s) { return s ? s->sim.getYDim() : 0; }
int ssb_get_zdim(ssb_sim* s) { return s ? s->sim.getZDim() : 0; }
double ssb_get_variable_at(ssb_sim* s, const char* id, int x, int y, int z) { return s && id ? s->sim.getVariableAt(id, x, y, z) : 0.0; }
void ssb_flip_volume_order(ssb_sim* s) { if (s) s->sim.flipVolumeOrder(); }

void ssb_release_variable(ssb_sim* s, const char* id) { if (s && id) s->sim.releaseVariable(id); }
double* ssb_get_variable(ssb_sim* s, const char* id, int* length) {
  int n = 0;
  double* p = s && id ? s->sim.getVariable(id, n) : NULL;
  if (length) *length = n;
  return p;
}
int* ssb_get_geometry(ssb_sim* s, const char* id, int* length) {
  int n = 0;
  int* p = s && id ? s->sim.getGeometry(id, n) : NULL;
  if (length) *length = n;
  return p;
}
uint8_t* ssb_get_boundary_type(ssb_sim* s, const char* id, int* length) {
  int n = 0;
  boundaryType* p = s && id ? s->sim.getBoundaryType(id, n) : NULL;
  if (length) *length = n;
  return reinterpret_cast<uint8_t*>(p);
}
int* ssb_get_boundary(ssb_sim* s, const char* id, int* length) {
  int n = 0;
  int* p = s && id ? s->sim.getBoundary(id, n) : NULL;
  if (length) *length = n;
  return p;
}
double* ssb_get_x(ssb_sim* s, int* length) { int n = 0; double* p = s ? s->sim.getX(n) : NULL; if (length) *length = n; return p; }
double* ssb_get_y(ssb_sim* s, int* length) { int n = 0; double* p = s ? s->sim.getY(n) : NULL; if (length) *length = n; return p; }
double* ssb_get_z(ssb_sim* s, int* length) { int n = 0; double* p = s ? s->sim.getZ(n) : NULL; if (length) *length = n; return p; }

int ssb_get_variable_compact(ssb_sim* s, const char* id, double* out, size_t n) { return s && id && s->sim.getVariableCompact(id, out, n) ? 0 : -1; }
int ssb_set_variable_compact(ssb_sim* s, const char* id, const double* in, size_t n) { return s && id && s->sim.setVariableCompact(id, in, n) ? 0 : -1; }
double ssb_get_stable_time_step(ssb_sim* s, double dt) {
  if (!s) return NAN;
  try { return s->sim.getStableTimeStep(dt); } catch (const std::exception& e) { s->error = e.what(); return NAN; }
}
long ssb_get_num_domain_cells(ssb_sim* s) { return s ? s->sim.getNumDomainCells() : 0; }
int ssb_get_num_staged_species(ssb_sim* s) { return s ? s->sim.getNumStagedSpecies() : 0; }
int ssb_get_time_slot(ssb_sim* s) { return s ? s->sim.getTimeSlot() : -1; }
long ssb_get_device_launch_count(ssb_sim* s) { return s ? s->sim.getDeviceLaunchCount() : 0; }
void ssb_device_state_changed(ssb_sim* s) { if (s) s->sim.deviceStateChanged(); }
void* ssb_get_device_handle(ssb_sim* s) { return s ? s->sim.getDeviceHandle() : NULL; }
int ssb_is_specialized(ssb_sim* s) { return s && s->sim.isSpecialized() ? 1 : 0; }
int ssb_specialize_now(ssb_sim* s) { return s && s->sim.specializeNow() ? 1 : 0; }
int ssb_get_specialize_info(ssb_sim* s, char* buf, int buflen) { return s ? copyText(s->sim.getSpecializeInfo(), buf, buflen) : 0; }
int ssb_get_device_program_text(ssb_sim* s, char* buf, int buflen) { return s ? copyText(s->sim.getDeviceProgramText(), buf, buflen) : 0; }
int ssb_get_setup_text(ssb_sim* s, const char* what, char* buf, int buflen) { return s && what ? copyText(s->sim.getSetupText(what), buf, buflen) : 0; }

}  // extern "C"
