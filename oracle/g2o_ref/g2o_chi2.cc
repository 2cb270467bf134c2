// Test infrastructure only: loads a .g2o text file with the REFERENCE'S OWN prebuilt g2o library
// (/root/reference/vido_slam/3rdparty/g2o/lib/libg2o.so), prints the chi2 its edge classes compute for it and optionally
// saves the graph again with the library's writer.  The g2o headers cannot be compiled here (no Eigen in the image), so the
// few members used are declared below with the library's names: the Itanium-mangled symbols are the library's
// (nm -DC libg2o.so: OptimizableGraph::load(char const*, bool), ::save(char const*, int) const, ::chi2() const,
// SparseOptimizer::SparseOptimizer(), ::initializeOptimization(int), ::computeActiveErrors(), ::activeChi2() const).
// The object is constructed by the library's constructor in a buffer far larger than the real class.
// usage: g2o_chi2 in.g2o [out.g2o]
#include <cstdio>
#include <cstdlib>
#include <new>

namespace g2o {
class OptimizableGraph {
 public:
  bool load(const char* filename, bool createEdges);
  bool save(const char* filename, int level) const;
  double chi2() const;
};
class SparseOptimizer : public OptimizableGraph {
 public:
  SparseOptimizer();
  bool initializeOptimization(int level);
  void computeActiveErrors();
  double activeChi2() const;
};
}  // namespace g2o

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s in.g2o [out.g2o]\n", argv[0]); return 2; }
  void* room = calloc(1, 1 << 20);
  g2o::SparseOptimizer* opt = new (room) g2o::SparseOptimizer();
  if (!opt->load(argv[1], true)) { fprintf(stderr, "load failed\n"); return 3; }
  if (!opt->initializeOptimization(0)) { fprintf(stderr, "initializeOptimization failed\n"); return 4; }
  opt->computeActiveErrors();
  printf("%.17g %.17g\n", opt->activeChi2(), opt->chi2());
  if (argc > 2 && !opt->save(argv[2], 0)) { fprintf(stderr, "save failed\n"); return 5; }
  fflush(stdout);
  _Exit(0);   // no destructor: the buffer is not the library's allocation
}
