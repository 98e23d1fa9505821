// delaunay.cu -- Delaunay triangulation of the support points on the GPU.
//
// Replaces computeDelaunayTriangulation (elas.cpp:445-505), i.e. Triangle 1.6
// run as triangulate("zQB") (triangle.cpp: divconqdelaunay 6160-6217,
// alternateaxes 5582-5601, divconqrecurse 5953-6103, mergehulls 5638-5934,
// writeelements 7800-7862).  Support points sit on a 5-px lattice, so
// cocircular quadruples are the rule and the Delaunay triangulation is not
// unique; the plane prior of the dense matcher depends on which diagonal is
// chosen.  To give the reference's triangles this kernel follows the same
// divide-and-conquer with alternating cuts and the same strict tie rules, but
// organised for a GPU:
//
//   * one CTA per (frame, image side); frames of a batch run concurrently;
//   * ordering: a stable LSD radix sort (7-bit digits, match.any ranking) of 27-bit (x,y) keys in
//     shared memory gives the (x,y) order and drops duplicates, two more passes the (y,x) order; the
//     alternating-axis median partition (a k-d tree build) is done level by
//     level with stable CTA-wide partitions of the two presorted lists;
//   * the recursion is unrolled into levels: all subproblems of one depth are
//     merged in parallel (one thread per merge), deepest level first;
//   * a subproblem of n points allocates exactly 2n-2 table rows in depth-first
//     order, so every thread knows its rows up front and the final row order
//     (= Triangle's allocation order = its output order) needs no atomics;
//   * the triangle table (3 neighbour handles + 3 vertex ids per row) lives in
//     shared memory as 6 x u16 per row (a handle = 4*row + orientation < 65536
//     for up to 16384 rows): the merges are serial pointer chasing, and a shared
//     memory access costs ~30 cycles against several hundred for L2;
//   * predicates are exact 64-bit integer determinants (coordinates < 2^13).
//
//   * point sets above SMEM_MAX_POINTS are cut at the D&C tree level whose subtrees fit: every
//     subtree is built in shared memory (all its levels), written out as 32-bit rows, and only the
//     few top-level merges (one, for up to 16 384 points) run on the table in global memory.
//
// merge_hulls() and leaf_case() below are a statement-for-statement transliteration of
// Triangle's mergehulls() / the base cases of divconqrecurse() onto table rows with integer
// predicates -- the same walk, the same strict comparisons, the same allocation order: the
// ORDERED triangle list must equal Triangle's, and that leaves no freedom in these two
// functions.  Everything around them (ordering, partition, level-synchronous scheduling, the
// table layouts, the subtree tiling) is this kernel's own organisation.
//
// Coincident points (right-image points (u-d,v) of two support points of a row; possible only with
// lr_threshold >= candidate_stepsize / 2, i.e. not with the presets): Triangle keeps the copy its
// randomised quicksort happens to put first.  The radix sort cannot know that order, so a point set
// that contains copies (and only such a set) has that quicksort replayed by one thread
// (vertexsort.cuh, checked against the compiled reference on the host) to pick the survivors.
#include "common.cuh"
#include "blockutil.cuh"
#include "vertexsort.cuh"

namespace {

constexpr int DT = 512;
constexpr int NW = DT / 32;
constexpr int SMEM_MAX_POINTS = 8192;    // one table in shared memory: (2n-2 rows * 12 B + n * 4 B) = 224 KB;
                                         // 2n-2 <= 16 382 rows keeps a handle (4*row + orientation) in 16 bits
constexpr int SORT_MAX = 16384;          // ordering + partition in shared memory (16-bit indices, 216 KB)
constexpr int RADIX_BITS = 7, RADIX = 1 << RADIX_BITS;
// dynamic shared memory, in 32 KB regions R0..R5, R6 (16 KB), R7 (8 KB) during ordering/partition
// (see delaunay_kernel), then the triangle table (192 KB) + packed coordinates (32 KB)
constexpr size_t DELAUNAY_SMEM = 224 * 1024;

struct Ot { int t, o; };

__device__ __forceinline__ int p1(int o) { return o == 2 ? 0 : o + 1; }
__device__ __forceinline__ int m1(int o) { return o == 0 ? 2 : o - 1; }
__device__ __forceinline__ int enc(Ot a) { return a.t * 4 + a.o; }
__device__ __forceinline__ Ot dec(int e) { Ot r; r.t = e >> 2; r.o = e & 3; return r; }
__device__ __forceinline__ Ot lnext(Ot a) { a.o = p1(a.o); return a; }
__device__ __forceinline__ Ot lprev(Ot a) { a.o = m1(a.o); return a; }

// Mesh vertex ids are POSITIONS in Triangle's final sortarray (0..nu-1; a subtree built in shared
// memory numbers its own points from 0); the emission maps them back to support indices.
// Triangle table in global memory, 32-bit entries.
struct MeshG {
  int* nb; int* vx; const int* x; const int* y;
  __device__ __forceinline__ int getnb(int t, int o) const { return nb[3 * t + o]; }
  __device__ __forceinline__ void setnb(int t, int o, int e) const { nb[3 * t + o] = e; }
  __device__ __forceinline__ int getvx(int t, int o) const { return vx[3 * t + o]; }
  __device__ __forceinline__ void setvx(int t, int o, int v) const { vx[3 * t + o] = v; }
  __device__ __forceinline__ int X(int v) const { return x[v]; }
  __device__ __forceinline__ int Y(int v) const { return y[v]; }
  __device__ __forceinline__ void XY(int v, int& px_, int& py_) const { px_ = x[v]; py_ = y[v]; }
};
// Triangle table in shared memory: row = {nb0,nb1,nb2,vx0,vx1,vx2} as u16; 0xFFFF = ghost vertex.
struct MeshS {
  unsigned short* rows; const unsigned* xy;   // xy[v] = x << 16 | y
  __device__ __forceinline__ int getnb(int t, int o) const { return rows[6 * t + o]; }
  __device__ __forceinline__ void setnb(int t, int o, int e) const { rows[6 * t + o] = (unsigned short)e; }
  __device__ __forceinline__ int getvx(int t, int o) const { int v = rows[6 * t + 3 + o]; return v == 0xFFFF ? -1 : v; }
  __device__ __forceinline__ void setvx(int t, int o, int v) const { rows[6 * t + 3 + o] = (unsigned short)v; }
  __device__ __forceinline__ int X(int v) const { return (int)(xy[v] >> 16); }
  __device__ __forceinline__ int Y(int v) const { return (int)(xy[v] & 0xFFFFu); }
  __device__ __forceinline__ void XY(int v, int& px_, int& py_) const {
    const unsigned c = xy[v];
    px_ = (int)(c >> 16);
    py_ = (int)(c & 0xFFFFu);
  }
};

template <class M> __device__ __forceinline__ Ot sym(const M& m, Ot a) { return dec(m.getnb(a.t, a.o)); }
template <class M> __device__ __forceinline__ int org(const M& m, Ot a) { return m.getvx(a.t, p1(a.o)); }
template <class M> __device__ __forceinline__ int dest(const M& m, Ot a) { return m.getvx(a.t, m1(a.o)); }
template <class M> __device__ __forceinline__ int apex(const M& m, Ot a) { return m.getvx(a.t, a.o); }
template <class M> __device__ __forceinline__ void setorg(const M& m, Ot a, int v) { m.setvx(a.t, p1(a.o), v); }
template <class M> __device__ __forceinline__ void setdest(const M& m, Ot a, int v) { m.setvx(a.t, m1(a.o), v); }
template <class M> __device__ __forceinline__ void setapex(const M& m, Ot a, int v) { m.setvx(a.t, a.o, v); }
template <class M> __device__ __forceinline__ void bond(const M& m, Ot a, Ot b) {
  m.setnb(a.t, a.o, enc(b));
  m.setnb(b.t, b.o, enc(a));
}
template <class M> __device__ __forceinline__ Ot newtri(const M& m, int row) {
  Ot r; r.t = row; r.o = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) { m.setnb(row, k, 0); m.setvx(row, k, -1); }
  return r;
}

// Exact predicates on coordinates.  |coordinate differences| < 2^13 + 2^12, so the orientation
// determinant fits in 32 bits; the incircle determinant needs 64 bits only for its last products.
struct Pt { int x, y; };
template <class M> __device__ __forceinline__ Pt load_pt(const M& m, int v) { Pt p; m.XY(v, p.x, p.y); return p; }
__device__ __forceinline__ int ccw_pt(Pt a, Pt b, Pt c) {
  return (a.x - c.x) * (b.y - c.y) - (a.y - c.y) * (b.x - c.x);
}
__device__ __forceinline__ bool incircle_pt(Pt a, Pt b, Pt c, Pt d) {
  const int adx = a.x - d.x, ady = a.y - d.y, bdx = b.x - d.x, bdy = b.y - d.y, cdx = c.x - d.x, cdy = c.y - d.y;
  const int al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;   // < 2^29
  const int t1 = bdx * cdy - cdx * bdy, t2 = cdx * ady - adx * cdy, t3 = adx * bdy - bdx * ady;   // < 2^29
  const long long det = (long long)al * t1 + (long long)bl * t2 + (long long)cl * t3;
  return det > 0;
}
template <class M> __device__ __forceinline__ long long ccw(const M& m, int a, int b, int c) {
  return ccw_pt(load_pt(m, a), load_pt(m, b), load_pt(m, c));
}

// Knit the triangulations of two adjacent point sets (mergehulls).  row0/row1 are
// the table rows of the bottom and top ghost this merge creates.
template <class M>
__device__ void merge_hulls(const M& m, Ot& farleft, Ot innerleft, Ot innerright, Ot& farright, int axis,
                            int row0, int row1) {
  int ild = dest(m, innerleft), ila = apex(m, innerleft);
  int iro = org(m, innerright), ira = apex(m, innerright);
  Ot check; int cv;
  if (axis == 1) {
    int flp = org(m, farleft), fla = apex(m, farleft);
    int frp = dest(m, farright);
    while (m.Y(fla) < m.Y(flp)) {
      farleft = sym(m, lnext(farleft));
      flp = fla;
      fla = apex(m, farleft);
    }
    check = sym(m, innerleft);
    cv = apex(m, check);
    while (m.Y(cv) > m.Y(ild)) {
      innerleft = lnext(check);
      ila = ild;
      ild = cv;
      check = sym(m, innerleft);
      cv = apex(m, check);
    }
    while (m.Y(ira) < m.Y(iro)) {
      innerright = sym(m, lnext(innerright));
      iro = ira;
      ira = apex(m, innerright);
    }
    check = sym(m, farright);
    cv = apex(m, check);
    while (m.Y(cv) > m.Y(frp)) {
      farright = lnext(check);
      frp = cv;
      check = sym(m, farright);
      cv = apex(m, check);
    }
  }
  bool changed;
  do {
    changed = false;
    if (ccw(m, ild, ila, iro) > 0) {
      innerleft = sym(m, lprev(innerleft));
      ild = ila;
      ila = apex(m, innerleft);
      changed = true;
    }
    if (ccw(m, ira, iro, ild) > 0) {
      innerright = sym(m, lnext(innerright));
      iro = ira;
      ira = apex(m, innerright);
      changed = true;
    }
  } while (changed);
  Ot leftcand = sym(m, innerleft), rightcand = sym(m, innerright);
  Ot base = newtri(m, row0);
  bond(m, base, innerleft);
  base = lnext(base);
  bond(m, base, innerright);
  base = lnext(base);
  setorg(m, base, iro);
  setdest(m, base, ild);
  if (ild == org(m, farleft)) farleft = lnext(base);
  if (iro == dest(m, farright)) farright = lprev(base);
  int ll = ild, lr = iro;
  int ul = apex(m, leftcand), ur = apex(m, rightcand);
  // coordinates of the four active vertices stay in registers
  Pt pll = load_pt(m, ll), plr = load_pt(m, lr), pul = load_pt(m, ul), pur = load_pt(m, ur);
  for (;;) {
    bool leftdone = ccw_pt(pul, pll, plr) <= 0;
    bool rightdone = ccw_pt(pur, pll, plr) <= 0;
    if (leftdone && rightdone) {
      Ot top = newtri(m, row1);
      setorg(m, top, ll);
      setdest(m, top, lr);
      bond(m, top, base);
      top = lnext(top);
      bond(m, top, rightcand);
      top = lnext(top);
      bond(m, top, leftcand);
      if (axis == 1) {
        int flp = org(m, farleft), frp = dest(m, farright), fra = apex(m, farright);
        check = sym(m, farleft);
        cv = apex(m, check);
        while (m.X(cv) < m.X(flp)) {
          farleft = lprev(check);
          flp = cv;
          check = sym(m, farleft);
          cv = apex(m, check);
        }
        while (m.X(fra) > m.X(frp)) {
          farright = sym(m, lprev(farright));
          frp = fra;
          fra = apex(m, farright);
        }
      }
      return;
    }
    if (!leftdone) {
      Ot nxt = sym(m, lprev(leftcand));
      int na = apex(m, nxt);
      if (na != -1) {
        Pt pna = load_pt(m, na);
        bool bad = incircle_pt(pll, plr, pul, pna);
        while (bad) {
          nxt = lnext(nxt);
          Ot topc = sym(m, nxt);
          nxt = lnext(nxt);
          Ot sidec = sym(m, nxt);
          bond(m, nxt, topc);
          bond(m, leftcand, sidec);
          leftcand = lnext(leftcand);
          Ot outerc = sym(m, leftcand);
          nxt = lprev(nxt);
          bond(m, nxt, outerc);
          setorg(m, leftcand, ll);
          setdest(m, leftcand, -1);
          setapex(m, leftcand, na);
          setorg(m, nxt, -1);
          setdest(m, nxt, ul);
          setapex(m, nxt, na);
          ul = na;
          pul = pna;
          nxt = sidec;
          na = apex(m, nxt);
          bad = false;
          if (na != -1) {
            pna = load_pt(m, na);
            bad = incircle_pt(pll, plr, pul, pna);
          }
        }
      }
    }
    if (!rightdone) {
      Ot nxt = sym(m, lnext(rightcand));
      int na = apex(m, nxt);
      if (na != -1) {
        Pt pna = load_pt(m, na);
        bool bad = incircle_pt(pll, plr, pur, pna);
        while (bad) {
          nxt = lprev(nxt);
          Ot topc = sym(m, nxt);
          nxt = lprev(nxt);
          Ot sidec = sym(m, nxt);
          bond(m, nxt, topc);
          bond(m, rightcand, sidec);
          rightcand = lprev(rightcand);
          Ot outerc = sym(m, rightcand);
          nxt = lnext(nxt);
          bond(m, nxt, outerc);
          setorg(m, rightcand, -1);
          setdest(m, rightcand, lr);
          setapex(m, rightcand, na);
          setorg(m, nxt, ur);
          setdest(m, nxt, -1);
          setapex(m, nxt, na);
          ur = na;
          pur = pna;
          nxt = sidec;
          na = apex(m, nxt);
          bad = false;
          if (na != -1) {
            pna = load_pt(m, na);
            bad = incircle_pt(pll, plr, pur, pna);
          }
        }
      }
    }
    if (leftdone || (!rightdone && incircle_pt(pul, pll, plr, pur))) {
      bond(m, base, rightcand);
      base = lprev(rightcand);
      setdest(m, base, ll);
      lr = ur;
      plr = pur;
      rightcand = sym(m, base);
      ur = apex(m, rightcand);
      pur = load_pt(m, ur);
    } else {
      bond(m, base, leftcand);
      base = lnext(leftcand);
      setorg(m, base, lr);
      ll = ul;
      pll = pul;
      leftcand = sym(m, base);
      ul = apex(m, leftcand);
      pul = load_pt(m, ul);
    }
  }
}

// Base cases of divconqrecurse: 2 points = an edge (2 ghosts), 3 points = a
// triangle + 3 ghosts or two edges (4 ghosts).  The points are vertex ids v0, v0+1[, v0+2].
template <class M>
__device__ void leaf_case(const M& m, int v0, int n, int row, Ot& farleft, Ot& farright) {
  const int sa[3] = {v0, v0 + 1, v0 + 2};
  if (n == 2) {
    Ot a = newtri(m, row);
    setorg(m, a, sa[0]);
    setdest(m, a, sa[1]);
    Ot b = newtri(m, row + 1);
    setorg(m, b, sa[1]);
    setdest(m, b, sa[0]);
    bond(m, a, b);
    a = lprev(a); b = lnext(b);
    bond(m, a, b);
    a = lprev(a); b = lnext(b);
    bond(m, a, b);
    farright = b;
    farleft = lprev(b);
    return;
  }
  Ot mid = newtri(m, row), t1 = newtri(m, row + 1), t2 = newtri(m, row + 2), t3 = newtri(m, row + 3);
  long long area = ccw(m, sa[0], sa[1], sa[2]);
  if (area == 0) {
    setorg(m, mid, sa[0]); setdest(m, mid, sa[1]);
    setorg(m, t1, sa[1]);  setdest(m, t1, sa[0]);
    setorg(m, t2, sa[2]);  setdest(m, t2, sa[1]);
    setorg(m, t3, sa[1]);  setdest(m, t3, sa[2]);
    bond(m, mid, t1);
    bond(m, t2, t3);
    mid = lnext(mid); t1 = lprev(t1); t2 = lnext(t2); t3 = lprev(t3);
    bond(m, mid, t3);
    bond(m, t1, t2);
    mid = lnext(mid); t1 = lprev(t1); t2 = lnext(t2); t3 = lprev(t3);
    bond(m, mid, t1);
    bond(m, t2, t3);
    farleft = t1;
    farright = t2;
  } else {
    setorg(m, mid, sa[0]);
    setdest(m, t1, sa[0]);
    setorg(m, t3, sa[0]);
    if (area > 0) {
      setdest(m, mid, sa[1]); setorg(m, t1, sa[1]); setdest(m, t2, sa[1]);
      setapex(m, mid, sa[2]); setorg(m, t2, sa[2]); setdest(m, t3, sa[2]);
    } else {
      setdest(m, mid, sa[2]); setorg(m, t1, sa[2]); setdest(m, t2, sa[2]);
      setapex(m, mid, sa[1]); setorg(m, t2, sa[1]); setdest(m, t3, sa[1]);
    }
    bond(m, mid, t1);
    mid = lnext(mid);
    bond(m, mid, t2);
    mid = lnext(mid);
    bond(m, mid, t3);
    t1 = lprev(t1); t2 = lnext(t2);
    bond(m, t1, t2);
    t1 = lprev(t1); t3 = lprev(t3);
    bond(m, t1, t3);
    t2 = lnext(t2); t3 = lprev(t3);
    bond(m, t2, t3);
    farleft = t1;
    farright = (area > 0) ? t2 : lnext(farleft);
  }
}

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Node (d,k) of the D&C tree over nu points: follow the bits of k from the root.  Returns false
// if the node does not exist (an ancestor was already a leaf).
__device__ __forceinline__ bool locate_node(int nu, int d, int k, int& lo, int& cnt, int& row) {
  lo = 0; cnt = nu; row = 0;
  for (int b = d - 1; b >= 0; b--) {
    if (cnt <= 3) return false;
    const int dv = cnt >> 1;
    if ((k >> b) & 1) { lo += dv; row += 2 * dv - 2; cnt -= dv; }
    else cnt = dv;
  }
  return true;
}

// All merges of the levels dfrom (deepest) .. dto of the subtree rooted at node (L,s), one thread
// per node, a CTA barrier between levels.  The mesh numbers rows and vertices relative to the
// subtree (row_off, pos_off); nodeL/nodeR carry the far-left / far-right handles in that numbering.
template <class M>
__device__ void build_levels(const M& m, int nu, int dfrom, int dto, int L, int s, int row_off, int pos_off,
                             int* nodeL, int* nodeR) {
  // Merges are serial pointer chasing with data-dependent control flow: threads of one warp that
  // run different merges serialise each other.  Nodes are therefore dealt to warps first (node j of a
  // round -> lane j / NW of warp j % NW): the top levels run one merge per warp.
  const int tid = threadIdx.x;
  const int spread = (tid & 31) * NW + (tid >> 5);
  for (int d = dfrom; d >= dto; d--) {
    const int sub = 1 << (d - L);
    for (int j = spread; j < sub; j += DT) {
      const int k = (s << (d - L)) + j;
      int lo, cnt, row;
      if (!locate_node(nu, d, k, lo, cnt, row)) continue;
      Ot fl, fr;
      if (cnt <= 3) {
        leaf_case(m, lo - pos_off, cnt, row - row_off, fl, fr);
      } else {
        const int c0 = (1 << (d + 1)) + 2 * k;
        fl = dec(nodeL[c0]);
        Ot il = dec(nodeR[c0]);
        Ot ir = dec(nodeL[c0 + 1]);
        fr = dec(nodeR[c0 + 1]);
        merge_hulls(m, fl, il, ir, fr, d & 1, row - row_off + 2 * cnt - 4, row - row_off + 2 * cnt - 3);
      }
      nodeL[(1 << d) + k] = enc(fl);
      nodeR[(1 << d) + k] = enc(fr);
    }
    __syncthreads();
#ifdef JN_DL_LEVEL_TIMES
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) printf("level %d (L %d s %d) done at %lld ns\n", d, L, s, gtime());
#endif
  }
}

// The non-ghost rows in row order (writeelements), vertex positions mapped back to support indices.
template <class M>
__device__ int emit_triangles(const M& m, const int* sa, int nu, int* flag, int* tri, int* s_part) {
  const int tid = threadIdx.x;
  const int rows = 2 * nu - 2;
  for (int t = tid; t < rows; t += DT)
    flag[t] = (m.getvx(t, 0) >= 0 && m.getvx(t, 1) >= 0 && m.getvx(t, 2) >= 0);
  __syncthreads();
  const int nt = block_exclusive_scan(flag, rows, s_part);
  for (int t = tid; t < rows; t += DT) {
    int v0 = m.getvx(t, 0), v1 = m.getvx(t, 1), v2 = m.getvx(t, 2);
    if (v0 >= 0 && v1 >= 0 && v2 >= 0) {
      int k = flag[t];
      tri[3 * k] = sa[v1];       // org
      tri[3 * k + 1] = sa[v2];   // dest
      tri[3 * k + 2] = sa[v0];   // apex
    }
  }
  return nt;
}

// One stable LSD radix pass over n (key, index) pairs in shared memory, 7-bit digit at `shift`.
// Warp w owns a contiguous slice; hist[digit * NW + w] counts its digits, an exclusive scan in
// (digit, warp) order turns the counts into output offsets, and the scatter ranks equal digits
// inside a 32-element step with match.any (earlier lanes first), so the pass is stable.
template <typename K>
__device__ void radix_pass(const K* key, const unsigned short* idx, K* okey, unsigned short* oidx, int n, int shift,
                           int* hist, int* s_part) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int slice = ((n + DT - 1) / DT) * 32;
  const int lo = min(w * slice, n), hi = min(lo + slice, n);
  for (int i = tid; i < RADIX * NW; i += DT) hist[i] = 0;
  __syncthreads();
  for (int i = lo + lane; i < hi; i += 32) atomicAdd(&hist[(((unsigned)key[i] >> shift) & (RADIX - 1)) * NW + w], 1);
  __syncthreads();
  block_exclusive_scan(hist, RADIX * NW, s_part);
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    const bool act = i < hi;
    K kv = 0;
    unsigned short iv = 0;
    if (act) { kv = key[i]; iv = idx[i]; }
    const unsigned d = act ? (((unsigned)kv >> shift) & (RADIX - 1)) : (unsigned)RADIX;
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    int old = 0;
    if (act && lane == leader) {
      old = hist[d * NW + w];
      hist[d * NW + w] = old + __popc(m);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    if (act) {
      const int pos = old + __popc(m & ((1u << lane) - 1u));
      okey[pos] = kv;
      oidx[pos] = iv;
    }
    __syncwarp();
  }
  __syncthreads();
}

// Alternating-axis median partition (alternateaxes), level by level: at every level each segment
// of >= 4 points is cut at its median along the level's axis; the list sorted along the other axis
// is split stably to match.  Returns the number of levels that split something; the final order
// (Triangle's sortarray) is left in xl.
template <typename TI, typename TF>
__device__ int partition_levels(TI*& xl, TI*& yl, TI*& sp, TI* scan, TI* seglo, TI* segn, TF* flag, int nu,
                                int* s_part, int* s_flag) {
  const int tid = threadIdx.x;
  for (int i = tid; i < nu; i += DT) { seglo[i] = 0; segn[i] = (TI)nu; }
  __syncthreads();
  int depth = 0;
  for (;; depth++) {
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    TI* prim = (depth & 1) ? yl : xl;
    TI* sec = (depth & 1) ? xl : yl;
    for (int i = tid; i < nu; i += DT) {
      int ns = segn[i];
      int r = 0;
      if (ns >= 4) { r = (i - (int)seglo[i]) >= (ns >> 1); *s_flag = 1; }
      flag[prim[i]] = (TF)r;
    }
    __syncthreads();
    if (!*s_flag) break;
    for (int i = tid; i < nu; i += DT) scan[i] = (TI)flag[sec[i]];
    __syncthreads();
    block_exclusive_scan(scan, nu, s_part);
    for (int i = tid; i < nu; i += DT) {
      int ns = segn[i], lo = seglo[i], id = sec[i];
      int pos = i;
      if (ns >= 4) {
        int rb = (int)scan[i] - (int)scan[lo];
        pos = flag[id] ? lo + (ns >> 1) + rb : lo + (i - lo - rb);
      }
      sp[pos] = (TI)id;
    }
    __syncthreads();
    for (int i = tid; i < nu; i += DT) {
      int ns = segn[i], lo = seglo[i];
      if (ns >= 4) {
        int dv = ns >> 1;
        if (i - lo < dv) segn[i] = (TI)dv;
        else { seglo[i] = (TI)(lo + dv); segn[i] = (TI)(ns - dv); }
      }
    }
    // the partitioned copy becomes the secondary list
    if (depth & 1) { TI* t = xl; xl = sp; sp = t; } else { TI* t = yl; yl = sp; sp = t; }
    __syncthreads();
  }
  return depth;
}


constexpr int OCC_EMPTY = 0x7f7f7f7f;   // cudaMemset(0x7f) pattern

// vertexsort_replay operands: support indices with their packed (x,y) keys in a table (shared
// memory path), or (key, index) records (global memory path)
struct KeyByIndex {
  const unsigned* k;
  __host__ __device__ unsigned operator()(unsigned short v) const { return k[v]; }
};
struct KeyVertex { unsigned k; int v; };
struct KeyOfRecord {
  __host__ __device__ unsigned operator()(const KeyVertex& r) const { return r.k; }
};

__global__ void __launch_bounds__(DT, 1)
delaunay_kernel(Geo g, Workspace ws) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ int s_part[DT + 1];
  __shared__ int s_flag;
  const int side = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
  FrameInfo* info = ws.info + frame;
  if (info->status != JN_OK) return;
  const int n = info->n_support;
  const size_t fo = (size_t)frame * g.cap_s;
  const int* px = ws.px[side] + fo;
  const int* py = ws.py + fo;
  int* const sa_g = ws.xlist[side] + fo;                        // final sortarray (support indices)
  int* const flag_g = ws.tmpD[side] + (size_t)frame * g.cap_t;  // cap_t ints
  int* const posx = ws.tmpB[side] + fo;                         // coordinates by sortarray position
  int* const posy = ws.tmpC[side] + fo;
  if (tid == 0) info->dt[side][0] = gtime();

  int nu, depth;
  const int* sa;
  if (n <= g.dl_sort_max) {
    // ---- 1+2 in shared memory, 16-bit indices.  Regions (bytes): R0 [0,32K) R1 [32K,64K) R2 [64K,96K)
    // R3 [96K,128K) R4 [128K,160K) R5 [160K,192K) R6 [192K,208K) R7 [208K,216K)
    typedef unsigned short u16;
    u16* const R0 = reinterpret_cast<u16*>(dsm);
    u16* const R1 = R0 + SORT_MAX;
    u16* const R2 = R1 + SORT_MAX;
    u16* const R3 = R2 + SORT_MAX;
    u16* const R4 = R3 + SORT_MAX;
    u16* const R5 = R4 + SORT_MAX;
    unsigned char* const R6 = reinterpret_cast<unsigned char*>(R5 + SORT_MAX);
    int* const hist = reinterpret_cast<int*>(R6 + SORT_MAX);
    // 1a. (x,y) order, stable in the support index: 4 passes over 27-bit keys x << 13 | y
    unsigned* K0 = reinterpret_cast<unsigned*>(R2);
    unsigned* K1 = reinterpret_cast<unsigned*>(R4);
    for (int i = tid; i < n; i += DT) {
      K0[i] = ((unsigned)px[i] << 13) | (unsigned)py[i];
      R0[i] = (u16)i;
    }
    __syncthreads();
    radix_pass(K0, R0, K1, R1, n, 0, hist, s_part);
    radix_pass(K1, R1, K0, R0, n, RADIX_BITS, hist, s_part);
    radix_pass(K0, R0, K1, R1, n, 2 * RADIX_BITS, hist, s_part);
    radix_pass(K1, R1, K0, R0, n, 3 * RADIX_BITS, hist, s_part);
    // 1b. the first of every (x,y) run survives (K1 is free: scan scratch)
    int* scan32 = reinterpret_cast<int*>(K1);
    for (int i = tid; i < n; i += DT) scan32[i] = (i == 0) || (K0[i] != K0[i - 1]);
    __syncthreads();
    nu = block_exclusive_scan(scan32, n, s_part);
    if (nu < n) {
      // Coincident points.  Which copy is "first" in Triangle's sortarray is decided by its randomised
      // quicksort (vertexsort.cuh): replay it -- serial, one thread, keys by support index in K1, the
      // index array in R1, the stack of pending parts in R6 -- and take the order of the copies inside
      // every run from it.  The runs themselves sit where the radix sort put them (same keys).
      unsigned* KI = K1;
      u16* A = R1;
      for (int i = tid; i < n; i += DT) {
        KI[i] = ((unsigned)px[i] << 13) | (unsigned)py[i];
        A[i] = (u16)i;
      }
      __syncthreads();
      if (tid == 0) {
        KeyByIndex key;
        key.k = KI;
        s_flag = vertexsort_replay(A, n, key, reinterpret_cast<u16*>(R6), SORT_MAX / 4) ? 1 : 0;
      }
      __syncthreads();
      if (!s_flag) {   // more than 4 096 pending parts: not a quicksort that terminates in this lifetime
        if (tid == 0) { info->status = JN_ERR_UNSUPPORTED; info->n_tri[side] = 0; }
        return;
      }
      for (int i = tid; i < n; i += DT) R0[i] = A[i];
      __syncthreads();
      for (int i = tid; i < n; i += DT) scan32[i] = (i == 0) || (K0[i] != K0[i - 1]);
      __syncthreads();
      nu = block_exclusive_scan(scan32, n, s_part);
    }
    u16* xl = R1;
    for (int i = tid; i < n; i += DT)
      if ((i == 0) || (K0[i] != K0[i - 1])) xl[scan32[i]] = R0[i];
    __syncthreads();
    if (nu < 2) {
      if (tid == 0) info->n_tri[side] = 0;
      return;
    }
    // 1c. (y,x) order: stable sort of the (x,y)-ordered list by y alone, 2 passes over 13-bit keys
    for (int i = tid; i < nu; i += DT) { R2[i] = (u16)py[xl[i]]; R4[i] = xl[i]; }
    __syncthreads();
    radix_pass(R2, R4, R3, R5, nu, 0, hist, s_part);
    radix_pass(R3, R5, R2, R4, nu, RADIX_BITS, hist, s_part);
    u16* yl = R4;
    if (tid == 0) info->dt[side][1] = gtime();
    // 2. partition
    u16* sp = R0;
    depth = partition_levels<u16, unsigned char>(xl, yl, sp, R2, R3, R5, R6, nu, s_part, &s_flag);
    for (int i = tid; i < nu; i += DT) {
      const int id = xl[i];
      sa_g[i] = id;
      posx[i] = px[id];
      posy[i] = py[id];
    }
    sa = sa_g;
  } else if (g.p.add_corners) {
    // corner points are off the lattice the occupancy grid is built on: not supported together
    if (tid == 0) { info->status = JN_ERR_UNSUPPORTED; info->n_tri[side] = 0; }
    return;
  } else {
    // ---- very large point sets: everything in global memory; ranks by prefix sums over an
    // occupancy grid (preset to OCC_EMPTY)
    int *xl = sa_g, *yl = ws.ylist[side] + fo, *sp = ws.tmpA[side] + fo;
    int* scan = ws.nodeL[side] + (size_t)frame * g.cap_t;         // reused before the merge phase
    const int step = g.p.candidate_stepsize;
    const int xdim = side ? g.W : g.Wc, xdiv = side ? 1 : step, Hc = g.Hc;
    const int cells = xdim * Hc;
    int* occ = ws.occ + ((size_t)frame * 2 + side) * ((size_t)g.W * Hc);
    int* scanbig = ws.trimap[side] + (size_t)frame * g.W * g.H;   // >= cells ints, free at this point
    const int xb = side ? JN_XBIAS : 0;
    for (int i = tid; i < n; i += DT) atomicMin(&occ[((px[i] - xb) / xdiv) * Hc + py[i] / step], i);
    __syncthreads();
    for (int c = tid; c < cells; c += DT) scanbig[c] = occ[c] != OCC_EMPTY;
    __syncthreads();
    nu = block_exclusive_scan(scanbig, cells, s_part);
    if (nu < n) {
      // Coincident points: the occupancy grid kept the lowest support index of every cell, Triangle
      // keeps the copy its quicksort puts first.  Replay the quicksort (vertexsort.cuh; serial, one
      // thread) on (key, index) records in the rows of the global triangle table, which the merges
      // only start to use later, with the stack in the idle shared memory, and put the first record of
      // every run into its cell.  Which cells are occupied does not change: the scan above stands.
      KeyVertex* A = reinterpret_cast<KeyVertex*>(ws.nb[side] + (size_t)frame * g.cap_t * 3);
      for (int i = tid; i < n; i += DT) {
        A[i].k = ((unsigned)px[i] << 13) | (unsigned)py[i];
        A[i].v = i;
      }
      __syncthreads();
      if (tid == 0)
        s_flag = vertexsort_replay(A, n, KeyOfRecord(), reinterpret_cast<int*>(dsm), (int)(DELAUNAY_SMEM / 8)) ? 1 : 0;
      __syncthreads();
      if (!s_flag) {
        if (tid == 0) { info->status = JN_ERR_UNSUPPORTED; info->n_tri[side] = 0; }
        return;
      }
      for (int i = tid; i < n; i += DT) {
        if (i == 0 || A[i].k != A[i - 1].k) {
          const int id = A[i].v;
          occ[((px[id] - xb) / xdiv) * Hc + py[id] / step] = id;
        }
      }
      __syncthreads();
    }
    for (int c = tid; c < cells; c += DT) {
      int id = occ[c];
      if (id != OCC_EMPTY) xl[scanbig[c]] = id;
    }
    __syncthreads();
    for (int j = tid; j < cells; j += DT) {   // y-major traversal: j = cy*xdim + cx
      int cy = j / xdim, cx = j - cy * xdim;
      scanbig[j] = occ[cx * Hc + cy] != OCC_EMPTY;
    }
    __syncthreads();
    block_exclusive_scan(scanbig, cells, s_part);
    for (int j = tid; j < cells; j += DT) {
      int cy = j / xdim, cx = j - cy * xdim;
      int id = occ[cx * Hc + cy];
      if (id != OCC_EMPTY) yl[scanbig[j]] = id;
    }
    __syncthreads();
    if (nu < 2) {
      if (tid == 0) info->n_tri[side] = 0;
      return;
    }
    if (tid == 0) info->dt[side][1] = gtime();
    // seglo/segn/flag of the partition live in the arrays that later hold the positions' coordinates
    // and the emission flags; the final order can end up in any of the three list buffers
    depth = partition_levels<int, int>(xl, yl, sp, scan, posx, posy, flag_g, nu, s_part, &s_flag);
    if (xl != sa_g) {
      for (int i = tid; i < nu; i += DT) sa_g[i] = xl[i];
    }
    __syncthreads();
    for (int i = tid; i < nu; i += DT) {
      const int id = sa_g[i];
      posx[i] = px[id];
      posy[i] = py[id];
    }
    sa = sa_g;
  }
  __syncthreads();   // the shared-memory lists are dead, sa_g / posx / posy are written
  if (tid == 0) { info->dt[side][2] = gtime(); info->dmerge_depth = depth; }

  // ---- 3. merges + emission --------------------------------------------------------------------
  int* nodeL = ws.nodeL[side] + (size_t)frame * g.cap_t;
  int* nodeR = ws.nodeR[side] + (size_t)frame * g.cap_t;
  int* tri = ws.tri[side] + (size_t)frame * g.cap_t * 3;
  MeshS ms;
  ms.rows = reinterpret_cast<unsigned short*>(dsm);
  unsigned* xy = reinterpret_cast<unsigned*>(dsm + (size_t)(2 * SMEM_MAX_POINTS) * 12);
  ms.xy = xy;
  MeshG mg;
  mg.nb = ws.nb[side] + (size_t)frame * g.cap_t * 3;
  mg.vx = ws.vx[side] + (size_t)frame * g.cap_t * 3;
  mg.x = posx; mg.y = posy;
  const int smem_max = min(g.dl_smem_max, SMEM_MAX_POINTS);
  int nt;
  if (nu <= smem_max) {
    for (int i = tid; i < nu; i += DT) xy[i] = ((unsigned)posx[i] << 16) | (unsigned)posy[i];
    __syncthreads();
    build_levels(ms, nu, depth, 0, 0, 0, 0, 0, nodeL, nodeR);
    nt = emit_triangles(ms, sa, nu, flag_g, tri, s_part);
  } else {
    // cut the tree at the first level L whose subtrees fit the shared-memory table (a node of c
    // points has children of c/2 and c - c/2 points); a table needs at least 2 points per subtree
    int L = 0;
    for (int c = nu; c > smem_max && L < depth; c -= c >> 1) L++;
    if (smem_max < 4) L = depth + 1;             // test hook: everything on the global table
    if (L <= depth) {
      for (int s = 0; s < (1 << L); s++) {
        int lo, cnt, row;
        if (!locate_node(nu, L, s, lo, cnt, row)) continue;
        for (int i = tid; i < cnt; i += DT) xy[i] = ((unsigned)posx[lo + i] << 16) | (unsigned)posy[lo + i];
        __syncthreads();
        build_levels(ms, nu, depth, L, L, s, row, lo, nodeL, nodeR);
        // write the subtree out: rows and vertex positions renumbered to the whole problem
        const int nrows = (cnt <= 3) ? ((cnt == 2) ? 2 : 4) : 2 * cnt - 2;
        for (int q = tid; q < nrows * 3; q += DT) {
          const int t = q / 3, o = q - 3 * t;
          mg.nb[3 * (row + t) + o] = ms.getnb(t, o) + 4 * row;
          const int v = ms.getvx(t, o);
          mg.vx[3 * (row + t) + o] = v < 0 ? -1 : v + lo;
        }
        if (tid == 0) {
          nodeL[(1 << L) + s] += 4 * row;
          nodeR[(1 << L) + s] += 4 * row;
        }
        __syncthreads();
      }
      if (tid == 0) info->dt[side][3] = gtime();
      build_levels(mg, nu, L - 1, 0, 0, 0, 0, 0, nodeL, nodeR);
    } else {
      build_levels(mg, nu, depth, 0, 0, 0, 0, 0, nodeL, nodeR);
    }
    nt = emit_triangles(mg, sa, nu, flag_g, tri, s_part);
  }
  if (tid == 0) { info->n_tri[side] = nt; info->dt[side][4] = gtime(); }
}

}  // namespace

// test hook: force the large-point-set paths (occupancy-grid ranking, global-memory tables)
static int g_sort_max = SORT_MAX, g_smem_max = SMEM_MAX_POINTS;
extern "C" void jn_debug_delaunay_limits(int sort_max, int smem_max) {
  g_sort_max = sort_max < 0 ? SORT_MAX : (sort_max < SORT_MAX ? sort_max : SORT_MAX);
  g_smem_max = smem_max < 0 ? SMEM_MAX_POINTS : (smem_max < SMEM_MAX_POINTS ? smem_max : SMEM_MAX_POINTS);
}

int launch_delaunay(const Geo& g_in, int B, Workspace& ws, cudaStream_t s) {
  Geo g = g_in;
  g.dl_sort_max = g_sort_max;
  g.dl_smem_max = g_smem_max;
  // per device (context), not per process: set on every launch, it is a host-side table write
  JN_CUDA_CHECK(cudaFuncSetAttribute(delaunay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)DELAUNAY_SMEM));
  // the occupancy grids (only used above the sort limit) start "empty" = 0x7f7f7f7f
  if (g.cap_s > g.dl_sort_max)
    JN_CUDA_CHECK(cudaMemsetAsync(ws.occ, 0x7f, (size_t)B * 2 * g.W * g.Hc * sizeof(int32_t), s));
  delaunay_kernel<<<dim3(2, B), DT, DELAUNAY_SMEM, s>>>(g, ws);
  g_jn_launches += 1;
  return JN_OK;
}
