// spatialsimulator.h — B200-backed SpatialSimulator with the public interface of the reference
// class (reference SpatialSBML/spatialsimulator.h:42-367): same method names, argument meaning
// and return conventions, so callers of that class can switch libraries.
//
// What differs behind the interface:
//   * all model set-up (SBML reading, geometry / membrane classification, boundary types,
//     AST → reverse-polish streams) runs on the host on COMPACT arrays (volume cells only);
//   * the time step runs on the GPU through the C ABI in include/spatial_b200.h;
//   * the doubled-grid arrays the reference hands out (getVariable, getGeometry, getX, ...) are
//     materialised on demand as host mirrors of the device state.
// There is no CPU implementation of the step: without a usable CUDA device init reports an error
// (isInitialized() == false, getLastError()) and oneStep() throws std::runtime_error.
#ifndef SPATIAL_SIMULATOR_B200_H
#define SPATIAL_SIMULATOR_B200_H

#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

class SBMLDocument;
class Model;

// Same layout as the reference's boundaryType (mystruct.h:28-35): six bools.
typedef struct _boundaryType {
  bool isBofXp;
  bool isBofXm;
  bool isBofYp;
  bool isBofYm;
  bool isBofZp;
  bool isBofZm;
} boundaryType;

struct SimImpl;  // host model + device binding (spatialsimulator.cpp)

class __attribute__((visibility("default"))) SpatialSimulator {
 public:
  SpatialSimulator();
  virtual ~SpatialSimulator();

  // reference spatialsimulator.h:66 / :78 — read SBML, set the model up, build the device program
  void initFromFile(const std::string& fileName, int xdim, int ydim, int zdim = 1);
  void initFromString(const std::string& sbml, int xdim, int ydim, int zdim = 1);

  // reference spatialsimulator.h:88 — one RK4 step; returns initialTime + step.
  // sim_time seen by the model is initialTime*step (reference spatialsimulator.cpp:1358).
  double oneStep(double initialTime, double step);

  void setParameterUniformly(const std::string& id, double value);  // :97
  void setParameter(const std::string& id, double value);           // :107
  int getIndexForPosition(double x, double y);                      // :116
  int getVariableLength() const;                                    // :123
  int getGeometryLength() const;                                    // :130
  int getXDim() const { return Xdiv; }
  int getYDim() const { return Ydiv; }
  int getZDim() const { return Zdiv; }
  double getVariableAt(const std::string& id, int x, int y, int z = 0);  // :157

  SpatialSimulator(SBMLDocument* doc, int xdim, int ydim, int zdim = 1);    // :165
  void initFromModel(SBMLDocument* doc, int xdim, int ydim, int zdim = 1);  // :177 (document stays owned by the caller)
  const Model* getModel() const;                                    // :182
  void flipVolumeOrder();                                           // :187

  // Borrowed pointers to doubled-grid host mirrors, valid until the simulator is re-initialised or
  // destroyed.  A variable pointer stays "live": values written through it are uploaded before the
  // next step, and the mirror is refreshed by the next accessor call after a step.
  double* getVariable(const std::string& id, int& length);            // :209
  // The pointer stays live like the reference's (which aliases the simulator's state): after every step the array is filled
  // again from the device, values written through it reach the device before the next step.  That costs a download per step
  // and handed-out variable; releaseVariable() ends it (the pointer stays valid, it is just no longer kept current).
  void releaseVariable(const std::string& id);
  int* getGeometry(const std::string& id, int& length);               // :220
  boundaryType* getBoundaryType(const std::string& id, int& length);  // :231
  int* getBoundary(const std::string& id, int& length);               // :242
  double* getX(int& length);                                          // :251
  double* getY(int& length);
  double* getZ(int& length);

  void setGnuplotExecutable(const std::string& location);  // :276
  const std::string& getGnuplotExecutable() const;
  void run(double endTime, double step);                   // :286 (no gnuplot pipe here)
  static int runOldMain(int argc, const char* argv[]);     // :291 (always -1, as in the reference)

  // ---- additions (not in the reference) ----
  bool isInitialized() const;
  const std::string& getLastError() const;
  // n steps with the step-counter time convention of run(): oneStep(firstStepIndex + i, step)
  void runSteps(double firstStepIndex, double step, int nsteps);
  // oneStep whose species values come from / go to caller-owned COMPACT host arrays (nx*ny*nz each, ideally pinned):
  // upload, the four stages and download are pipelined over slabs (sb2_step_host).  ids[i] names the species of
  // in[i] / out[i] (either may be NULL).  Falls back to upload + oneStep + download for models off the fused schedule.
  double oneStepHost(double initialTime, double step, int n, const char* const* ids, const double* const* in, double* const* out, int nslabs);
  // compact (volume cells only, nx*ny*nz) copy of a species / field, without the doubled mirror
  bool getVariableCompact(const std::string& id, double* out, size_t n);
  bool setVariableCompact(const std::string& id, const double* in, size_t n);
  // largest admissible dt by the reference's stability criteria (checkStability.cpp)
  double getStableTimeStep(double dt);
  long getNumDomainCells() const;  // volume cells that carry at least one staged species
  int getNumStagedSpecies() const;
  int getTimeSlot() const;  // constants-table slot of the `time` symbol (sb2_begin_step), -1: unused
  long getDeviceLaunchCount() const;
  // the caller stepped the device simulator itself (split-phase sb2_begin_step / sb2_stage / sb2_end_step of the
  // slab-sharded path): the host mirrors of the species fields are out of date
  void deviceStateChanged();
  void setCudaStream(void* stream);
  void setDevice(int ordinal);  // before init
  // Several GPUs of one box behind this one object: before init, name the devices; the grid is then cut into
  // one slab per entry along its slowest axis (y in 2-D, z in 3-D), every slab lives on its device (an ordinal may repeat),
  // and oneStep() runs the sharded step of the library (sb2_group_step: edge rows first, halo rows by peer copies, interior
  // rows meanwhile).  Results are bit-identical to a single device.  getVariableCompact / setVariableCompact / getVariable /
  // getVariableAt / setParameter work on the whole grid as before.
  void setDevices(const int* ordinals, int n);
  int getNumSlabs() const;
  // slab (setSlab): first row held locally and the owned rows [ownLo, ownHi), all in rows of the whole grid; zeros otherwise
  void getSlabRows(int& heldLo, int& ownLo, int& ownHi) const;
  // Slab of a larger grid (multi-GPU): this simulator owns rows [lo, hi) of the slab
  // axis (y in 2-D, z in 3-D) of the model's grid.  Call before init; lo = hi = 0 means everything.
  void setSlab(int lo, int hi);
  // Set the model up on the host only (geometry, masks, programs can be inspected; no GPU needed).
  // oneStep() then throws.
  void setHostOnly(bool hostOnly);
  std::string getDeviceProgramText();
  // model-specialised stage kernel (include/spatial_b200.h, sb2_specialized): whether it runs, and what happened at set-up
  bool isSpecialized();
  bool specializeNow();  // build it (or wait for the background build) before returning
  std::string getSpecializeInfo();
  std::string getSetupText(const std::string& what);  // "rxn/<i>", "geo/<compartment>", "summary"
  void* getDeviceHandle();  // sb2_sim*
  SimImpl* getImpl() { return impl; }

 private:
  SpatialSimulator(const SpatialSimulator&);
  SpatialSimulator& operator=(const SpatialSimulator&);
  void clear();
  bool buildDevice();
  bool buildSlabs(SBMLDocument* doc, int xdim, int ydim, int zdim);
  bool gatherCompact(const std::string& id, double* out);
  bool addMemTransport(void* rxn);
  bool addMemTransportOn(void* rxn, void* membraneGeo);
  bool addMembraneSpecies();
  bool addMembraneReaction(void* rxn, void* geo, int kind);
  void syncBeforeStep();
  void refreshHandedOut();

  int Xdiv, Ydiv, Zdiv;
  SimImpl* impl;
  std::string gnuplotExecutable;
  std::string lastError;
  int deviceOrdinal;
  void* externalStream;
  bool hostOnly;
  int slabLo, slabHi;
  std::vector<int> groupDevices;
};

#endif  // SPATIAL_SIMULATOR_B200_H
