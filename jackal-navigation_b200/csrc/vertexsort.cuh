// vertexsort.cuh -- which of several coincident points Triangle keeps.
//
// divconqdelaunay (triangle.cpp:6160-6217) orders the vertices with vertexsort
// (triangle.cpp:5446-5499), a quicksort whose pivots come from randomnation
// (triangle.cpp:4045-4049, seed reset to 1 by triangleinit, :4030), and then keeps the FIRST of
// every run of equal (x,y) (:6179-6195).  For distinct points the sorted order does not depend on
// the pivots, which is why the kernel sorts with a radix sort.  For coincident points it does:
// which copy ends up first is decided by the whole history of Hoare swaps.  The copies differ in
// everything but (x,y) -- right-image points (u-d,v) of different (u,d) -- so the survivor decides
// the vertex ids of the triangles around it and through them the right-image planes.
//
// vertexsort_replay() re-runs that quicksort as Triangle does, statement by statement (the same
// pivot sequence, the same scan and swap order, the left part before the right one), on an index
// array.  It is a serial algorithm: the delaunay kernel calls it from ONE thread and only for point
// sets that actually contain coincident points (none with the default parameters: candidates sit
// 5 px apart and the cross check allows a difference of 2).  Plain C++ so that the CPU tests can
// compile the very same function with g++ and compare it with the compiled reference
// (tests/test_host_logic.py).
#ifndef JN_VERTEXSORT_CUH
#define JN_VERTEXSORT_CUH

#ifdef __CUDACC__
#define JN_VS_FN __host__ __device__ inline
#else
#define JN_VS_FN inline
#endif

// randomnation, triangle.cpp:4045-4049.  seed < 714025, so seed * 1366 + 150889 < 2^30.
JN_VS_FN unsigned vs_randomnation(unsigned& seed, unsigned choices) {
  seed = (seed * 1366u + 150889u) % 714025u;
  return seed / (714025u / choices + 1u);
}

// a[0..n): the vertices (vertex numbers, or records that carry their key), in input order on entry
// and in Triangle's sorted order on return.
// key(a[i]): an unsigned value that orders vertices like (x, then y), e.g. x << 13 | y.
// stack: room for stack_cap (start, count) pairs of pending right-hand parts (one per level of the
// recursion: ~2 log2 n expected).  Returns false if the stack is too small (a is then unsorted).
template <class IDX, class STK, class KEY>
JN_VS_FN bool vertexsort_replay(IDX* a, int n, KEY key, STK* stack, int stack_cap) {
  unsigned seed = 1u;
  int sp = 0;
  if (n < 2) return true;
  if (stack_cap < 1) return false;
  stack[0] = (STK)0;
  stack[1] = (STK)n;
  sp = 1;
  while (sp > 0) {
    sp--;
    const int lo = (int)stack[2 * sp], cnt = (int)stack[2 * sp + 1];
    IDX* s = a + lo;
    if (cnt == 2) {   // triangle.cpp:5461-5470
      if (key(s[0]) > key(s[1])) { const IDX t = s[1]; s[1] = s[0]; s[0] = t; }
      continue;
    }
    const int pivot = (int)vs_randomnation(seed, (unsigned)cnt);
    const unsigned pk = key(s[pivot]);
    int left = -1, right = cnt;
    while (left < right) {   // triangle.cpp:5477-5497
      do { left++; } while (left <= right && key(s[left]) < pk);
      do { right--; } while (left <= right && key(s[right]) > pk);
      if (left < right) { const IDX t = s[left]; s[left] = s[right]; s[right] = t; }
    }
    // Triangle recurses into the left part first: the right part waits on the stack
    if (right < cnt - 2) {
      if (sp >= stack_cap) return false;
      stack[2 * sp] = (STK)(lo + right + 1);
      stack[2 * sp + 1] = (STK)(cnt - right - 1);
      sp++;
    }
    if (left > 1) {
      if (sp >= stack_cap) return false;
      stack[2 * sp] = (STK)lo;
      stack[2 * sp + 1] = (STK)left;
      sp++;
    }
  }
  return true;
}

#endif  // JN_VERTEXSORT_CUH
