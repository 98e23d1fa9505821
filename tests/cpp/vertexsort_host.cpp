// Host build of csrc/vertexsort.cuh (the function the delaunay kernel runs on the device for point
// sets with coincident points), for tests/test_host_logic.py: g++ compiles the same source.
#include <cstddef>
#include <cstdint>
#include <vector>
#include "vertexsort.cuh"

struct PackedKey {
  const int32_t* x; const int32_t* y;
  unsigned operator()(uint16_t v) const { return ((unsigned)x[v] << 13) | (unsigned)y[v]; }
};

// order[0..n) = vertex numbers in Triangle's sorted order; returns 0, or -1 if the stack was too small
extern "C" int vs_sorted_order(const int32_t* x, const int32_t* y, int n, int32_t* order, int stack_cap) {
  std::vector<uint16_t> a(n), st(2 * (std::size_t)(stack_cap > 0 ? stack_cap : 1));
  for (int i = 0; i < n; i++) a[i] = (uint16_t)i;
  PackedKey k{x, y};
  if (!vertexsort_replay(a.data(), n, k, st.data(), stack_cap)) return -1;
  for (int i = 0; i < n; i++) order[i] = a[i];
  return 0;
}

// the form the kernel uses for point sets above its shared-memory limit: (key, vertex) records
struct KeyVertex { unsigned k; int32_t v; };
struct KeyOfRecord { unsigned operator()(const KeyVertex& r) const { return r.k; } };
extern "C" int vs_sorted_order_records(const int32_t* x, const int32_t* y, int n, int32_t* order, int stack_cap) {
  std::vector<KeyVertex> a(n);
  std::vector<int> st(2 * (std::size_t)(stack_cap > 0 ? stack_cap : 1));
  for (int i = 0; i < n; i++) { a[i].k = ((unsigned)x[i] << 13) | (unsigned)y[i]; a[i].v = i; }
  if (!vertexsort_replay(a.data(), n, KeyOfRecord(), st.data(), stack_cap)) return -1;
  for (int i = 0; i < n; i++) order[i] = a[i].v;
  return 0;
}
