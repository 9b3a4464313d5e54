// ORACLE / TEST INFRASTRUCTURE — never linked into or called by the product library.
//
// Drives the reference's own, unmodified SpatialSimulator (compiled from /root/reference/SpatialSBML
// by oracle/Makefile against oracle/sbml_shim) and either dumps its state or times its oneStep
// loop.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs execute the built
// binary (oracle/_ref/ref_driver).
//
//   ref_driver dump <model.xml> <xdim> <ydim> <zdim> <dt> <out.sbd> [step ...]
//       state after init (masks, index lists, normals, RPN streams, fields) and species fields
//       after each listed cumulative step count; stepping follows SpatialSimulator::run
//       (reference spatialsimulator.cpp:1682-1694): oneStep(stepIndex, dt).
//   ref_driver time <model.xml> <xdim> <ydim> <zdim> <dt> <warmup> <steps>
//       prints one JSON line with wall seconds for <steps> oneStep calls.
//
// Container format "SBD1": repeated records  u32 nameLen, name, u8 dtype ('i' int32, 'd' float64,
// 'b' uint8, 's' text), u64 count, payload.
#define private public  // the dump needs geoInfoList / varInfoList / nuVec (test-only access)
#include "spatialsimulator.h"
#undef private

#include <sbml/SBMLTypes.h>
#include <malloc.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>
#include <vector>

static FILE* g_out = NULL;

static void rec(const std::string& name, char dtype, const void* data, uint64_t count, size_t elem) {
  uint32_t nl = (uint32_t)name.size();
  fwrite(&nl, 4, 1, g_out);
  fwrite(name.data(), 1, nl, g_out);
  fwrite(&dtype, 1, 1, g_out);
  fwrite(&count, 8, 1, g_out);
  if (count) fwrite(data, elem, count, g_out);
}
static void recI(const std::string& n, const std::vector<int>& v) { rec(n, 'i', v.data(), v.size(), 4); }
static void recD(const std::string& n, const double* v, uint64_t c) { rec(n, 'd', v, c, 8); }
static void recS(const std::string& n, const std::string& s) { rec(n, 's', s.data(), s.size(), 1); }

static const char* varNameOf(SpatialSimulator& sim, const double* valuePtr) {
  for (size_t i = 0; i < sim.varInfoList.size(); ++i)
    if (sim.varInfoList[i]->value == valuePtr) return sim.varInfoList[i]->id;
  return NULL;
}

static std::string rpnText(SpatialSimulator& sim, reversePolishInfo* rp) {
  std::ostringstream os;
  os.precision(17);
  for (int i = 0; i < rp->listNum; ++i) {
    if (rp->varList && rp->varList[i]) {
      const char* nm = varNameOf(sim, rp->varList[i]);
      os << "V " << (nm ? nm : "?") << " " << ((rp->deltaList && rp->deltaList[i]) ? 1 : 0) << "\n";
    } else if (rp->constList && rp->constList[i]) {
      const char* nm = varNameOf(sim, rp->constList[i]);
      os << "C " << (nm ? nm : "#") << " " << *rp->constList[i] << "\n";
    } else {
      os << "O " << rp->opfuncList[i] << "\n";
    }
  }
  return os.str();
}

static void dumpFields(SpatialSimulator& sim, const std::string& prefix, int N) {
  for (unsigned int i = 0; i < sim.numOfSpecies; ++i) {
    const char* id = sim.model->getSpecies(i)->getId().c_str();
    variableInfo* info = NULL;
    for (size_t k = 0; k < sim.varInfoList.size(); ++k)
      if (strcmp(sim.varInfoList[k]->id, id) == 0) { info = sim.varInfoList[k]; break; }
    if (!info || !info->value) continue;
    if (info->isUniform) recD(prefix + id, info->value, 1);
    else recD(prefix + id, info->value, N);
  }
}

static int doDump(int argc, char** argv) {
  if (argc < 8) return 2;
  const char* file = argv[2];
  int xd = atoi(argv[3]), yd = atoi(argv[4]), zd = atoi(argv[5]);
  double dt = atof(argv[6]);
  g_out = fopen(argv[7], "wb");
  if (!g_out) { perror("open out"); return 1; }
  fwrite("SBD1", 1, 4, g_out);

  static SpatialSimulator sim;  // zero-initialised storage: the reference's constructor leaves sim_time unset, and rules read it at set-up
  sim.initFromFile(file, xd, yd, zd);
  int N = sim.numOfVolIndexes;

  std::vector<int> dims = {sim.Xdiv, sim.Ydiv, sim.Zdiv, sim.Xindex, sim.Yindex, sim.Zindex, (int)sim.dimension, N,
                           (int)sim.volDimension, (int)sim.memDimension};
  recI("dims", dims);
  double del[3] = {sim.deltaX, sim.deltaY, sim.deltaZ};
  recD("delta", del, 3);

  std::string geoNames;
  for (size_t g = 0; g < sim.geoInfoList.size(); ++g) {
    GeometryInfo* gi = sim.geoInfoList[g];
    std::string p = std::string("geo/") + gi->compartmentId + "/";
    geoNames += std::string(gi->compartmentId) + "\n";
    std::vector<int> flags = {gi->isVol ? 1 : 0};
    recI(p + "isVol", flags);
    recS(p + "domainType", gi->domainTypeId ? gi->domainTypeId : "");
    recS(p + "adjacent1", gi->adjacentGeo1 ? gi->adjacentGeo1->compartmentId : "");
    recS(p + "adjacent2", gi->adjacentGeo2 ? gi->adjacentGeo2->compartmentId : "");
    if (gi->isDomain) rec(p + "isDomain", 'i', gi->isDomain, N, 4);
    if (gi->isBoundary) rec(p + "isBoundary", 'i', gi->isBoundary, N, 4);
    if (gi->bType) rec(p + "bType", 'b', gi->bType, (uint64_t)N * sizeof(boundaryType), 1);
    recI(p + "domainIndex", gi->domainIndex);
    recI(p + "pseudoMemIndex", gi->pseudoMemIndex);
    recI(p + "boundaryIndex", gi->boundaryIndex);
    if (!gi->isVol && sim.nuVec) {
      std::vector<double> nv;
      for (size_t k = 0; k < gi->domainIndex.size(); ++k) {
        int idx = gi->domainIndex[k];
        nv.push_back(sim.nuVec[idx].nx); nv.push_back(sim.nuVec[idx].ny); nv.push_back(sim.nuVec[idx].nz);
      }
      recD(p + "nuVec", nv.data(), nv.size());
    }
    if (!gi->isVol && sim.vorI) {
      std::vector<double> vd; std::vector<int> vi;
      for (size_t k = 0; k < gi->domainIndex.size(); ++k) {
        const voronoiInfo& v = sim.vorI[gi->domainIndex[k]];
        for (int j = 0; j < 2; ++j) { vi.push_back(v.adjacentIndexXY[j]); vi.push_back(v.adjacentIndexYZ[j]); vi.push_back(v.adjacentIndexXZ[j]); }
        for (int j = 0; j < 2; ++j) { vd.push_back(v.siXY[j]); vd.push_back(v.siYZ[j]); vd.push_back(v.siXZ[j]); }
        for (int j = 0; j < 2; ++j) { vd.push_back(v.diXY[j]); vd.push_back(v.diYZ[j]); vd.push_back(v.diXZ[j]); }
      }
      recI(p + "vorAdj", vi);
      recD(p + "vorSiDi", vd.data(), vd.size());
    }
  }
  recS("geoNames", geoNames);

  std::string varNames;
  for (size_t k = 0; k < sim.varInfoList.size(); ++k) {
    variableInfo* vi = sim.varInfoList[k];
    varNames += std::string(vi->id) + "\n";
    std::string p = std::string("var/") + vi->id + "/";
    std::vector<int> flags = {vi->isUniform ? 1 : 0, vi->isResolved ? 1 : 0, vi->inVol ? 1 : 0, vi->delta ? 1 : 0,
                              vi->hasAssignmentRule ? 1 : 0, vi->value ? 1 : 0};
    recI(p + "flags", flags);
    recS(p + "geo", vi->geoi ? vi->geoi->compartmentId : "");
    if (vi->value) recD(p + "value", vi->value, vi->isUniform ? 1 : N);
  }
  recS("varNames", varNames);

  for (size_t r = 0; r < sim.rInfoList.size(); ++r) {
    reactionInfo* ri = sim.rInfoList[r];
    std::ostringstream p; p << "rxn/" << r << "/";
    recS(p.str() + "id", ri->id ? ri->id : "");
    recS(p.str() + "rpn", rpnText(sim, ri->rpInfo));
    std::vector<int> flags = {ri->isMemTransport ? 1 : 0, ri->reaction ? (int)ri->reaction->getNumReactants() : -1};
    recI(p.str() + "flags", flags);
    std::ostringstream refs; refs.precision(17);
    for (size_t k = 0; k < ri->spRefList.size(); ++k)
      refs << (ri->spRefList[k] ? ri->spRefList[k]->id : "?") << " " << ri->srStoichiometry[k] << " "
           << (ri->isVariable[k] ? 1 : 0) << "\n";
    recS(p.str() + "refs", refs.str());
  }
  for (size_t g = 0; g < sim.geoInfoList.size(); ++g)
    if (sim.geoInfoList[g]->rpInfo)
      recS(std::string("geo/") + sim.geoInfoList[g]->compartmentId + "/rpn", rpnText(sim, sim.geoInfoList[g]->rpInfo));

  dumpFields(sim, "step/0/", N);
  int done = 0;
  for (int a = 8; a < argc; ++a) {
    int target = atoi(argv[a]);
    for (; done < target; ++done) sim.oneStep(done, dt);
    std::ostringstream p; p << "step/" << target << "/";
    dumpFields(sim, p.str(), N);
  }
  fclose(g_out);
  return 0;
}

static int doTime(int argc, char** argv) {
  if (argc < 9) return 2;
  const char* file = argv[2];
  int xd = atoi(argv[3]), yd = atoi(argv[4]), zd = atoi(argv[5]);
  double dt = atof(argv[6]);
  int warm = atoi(argv[7]), steps = atoi(argv[8]);
  auto t0 = std::chrono::steady_clock::now();
  static SpatialSimulator sim;  // zero-initialised storage: the reference's constructor leaves sim_time unset, and rules read it at set-up
  sim.initFromFile(file, xd, yd, zd);
  auto t1 = std::chrono::steady_clock::now();
  int t = 0;
  for (; t < warm; ++t) sim.oneStep(t, dt);
  auto t2 = std::chrono::steady_clock::now();
  for (int k = 0; k < steps; ++k, ++t) sim.oneStep(t, dt);
  auto t3 = std::chrono::steady_clock::now();
  // cells = volume cells that host at least one non-uniform species (max over compartments summed)
  long cells = 0;
  for (size_t g = 0; g < sim.geoInfoList.size(); ++g) {
    GeometryInfo* gi = sim.geoInfoList[g];
    if (!gi->isVol) continue;
    bool used = false;
    for (size_t k = 0; k < sim.varInfoList.size(); ++k)
      if (sim.varInfoList[k]->sp && !sim.varInfoList[k]->isUniform && sim.varInfoList[k]->geoi == gi) used = true;
    if (used) cells += (long)gi->domainIndex.size();
  }
  double init_s = std::chrono::duration<double>(t1 - t0).count();
  double step_s = std::chrono::duration<double>(t3 - t2).count();
  printf("{\"init_s\": %.6f, \"steps\": %d, \"seconds\": %.6f, \"cells\": %ld, \"cell_updates_per_s\": %.6e}\n",
         init_s, steps, step_s, cells, step_s > 0 ? (double)cells * steps / step_s : 0.0);
  return 0;
}

// ref_driver stab <model.xml> <xdim> <ydim> <zdim> <dt>...
// The reference's stability reductions (checkStability.cpp:10-52, 119-154) are not wired into its public API; the
// oracle calls them directly for every species that diffuses / is advected and prints the smallest admissible step.
double checkDiffusionStab(variableInfo* sInfo, double deltaX, double deltaY, double deltaZ, int Xindex, int Yindex, double dt);
double checkAdvectionStab(variableInfo* sInfo, double deltaX, double deltaY, double deltaZ, double dt, int Xindex, int Yindex, int dimension);
double checkMemDiffusionStab(variableInfo* sInfo, voronoiInfo* vorI, int Xindex, int Yindex, double dt, double dimension);
static int doStab(int argc, char** argv) {
  if (argc < 7) return 2;
  static SpatialSimulator sim;  // zero-initialised storage: the reference's constructor leaves sim_time unset, and rules read it at set-up
  sim.initFromFile(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  printf("{");
  for (int a = 6; a < argc; ++a) {
    const double dt = atof(argv[a]);
    double best = dt;
    for (size_t k = 0; k < sim.varInfoList.size(); ++k) {
      variableInfo* v = sim.varInfoList[k];
      if (!v->sp || v->isUniform || !v->geoi) continue;
      // volume species: checkDiffusionStab, species on a membrane: checkMemDiffusionStab (the dispatch of the reference's own,
      // commented-out, call site spatialsimulator.cpp:1172-1179)
      if (v->diffCInfo && v->geoi->isVol) best = std::min(best, checkDiffusionStab(v, sim.deltaX, sim.deltaY, sim.deltaZ, sim.Xindex, sim.Yindex, dt));
      if (v->diffCInfo && !v->geoi->isVol && sim.vorI) best = std::min(best, checkMemDiffusionStab(v, sim.vorI, sim.Xindex, sim.Yindex, dt, (double)sim.dimension));
      if (v->adCInfo) best = std::min(best, checkAdvectionStab(v, sim.deltaX, sim.deltaY, sim.deltaZ, dt, sim.Xindex, sim.Yindex, (int)sim.dimension));
    }
    printf("%s\"%s\": %.17g", a > 6 ? ", " : "", argv[a], best);
  }
  printf("}\n");
  return 0;
}

int main(int argc, char** argv) {
  // The reference's membrane scan reads isDomain[index - Xindex*Yindex] on the Z = 0 plane of a 3-D grid (and the like on the
  // other frame planes, spatialsimulator.cpp:929-960): a few KB before the array.  With glibc's default mmap threshold such an
  // array is its own mapping and the read can fault, depending on the heap layout; keeping big arrays on the brk heap makes
  // the read land in mapped memory, as it does in a typical run of the reference.
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  if (argc >= 2 && std::string(argv[1]) == "stab") return doStab(argc, argv);
  if (argc >= 2 && std::string(argv[1]) == "dump") return doDump(argc, argv);
  if (argc >= 2 && std::string(argv[1]) == "time") return doTime(argc, argv);
  fprintf(stderr, "usage: ref_driver dump|time ... (see header comment)\n");
  return 2;
}
