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
//   * ordering: a bitonic sort of 64-bit (coordinate, coordinate, index) keys in
//     shared memory gives the (x,y) and (y,x) orders and drops duplicates; the
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
//     memory access costs ~30 cycles against several hundred for L2.  Point
//     sets too large for that (> SMEM_MAX_POINTS) use the same code on 32-bit
//     tables in global memory;
//   * predicates are exact 64-bit integer determinants (coordinates < 2^13).
//
// Known deviation: among duplicate right-image points (u-d,v) Triangle keeps
// the one its randomized quicksort happens to put first; this kernel keeps the
// lowest support index.
#include "common.cuh"
#include "blockutil.cuh"

namespace {

constexpr int DT = 512;
constexpr int SMEM_MAX_POINTS = 7680;                      // 2n-2 rows * 12 B + n * 4 B <= ~210 KB
constexpr int SORT_MAX = 8192;                             // bitonic sort capacity (64 KB of keys)
// dynamic shared memory: sort keys (64 KB), then the partition lists (216 KB), then the triangle
// table + packed coordinates (2*7680 rows * 12 B + 7680 * 4 B = 210 KB)
constexpr size_t DELAUNAY_SMEM = 220 * 1024;

struct Ot { int t, o; };

__device__ __forceinline__ int p1(int o) { return o == 2 ? 0 : o + 1; }
__device__ __forceinline__ int m1(int o) { return o == 0 ? 2 : o - 1; }
__device__ __forceinline__ int enc(Ot a) { return a.t * 4 + a.o; }
__device__ __forceinline__ Ot dec(int e) { Ot r; r.t = e >> 2; r.o = e & 3; return r; }
__device__ __forceinline__ Ot lnext(Ot a) { a.o = p1(a.o); return a; }
__device__ __forceinline__ Ot lprev(Ot a) { a.o = m1(a.o); return a; }

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
// triangle + 3 ghosts or two edges (4 ghosts).
template <class M>
__device__ void leaf_case(const M& m, const int* sa, int n, int row, Ot& farleft, Ot& farright) {
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

// All merges of one triangulation, deepest level first, then the non-ghost rows in row
// order (writeelements).  Returns the triangle count.
template <class M>
__device__ int build_and_emit(const M& m, const int* sa, int nu, int depth, int* nodeL, int* nodeR, int* flag,
                              int* tri, int* s_part) {
  const int tid = threadIdx.x;
  for (int d = depth; d >= 0; d--) {
    const int nodes = 1 << d;
    for (int k = tid; k < nodes; k += DT) {
      // locate node (d,k): follow the bits of k from the root
      int lo = 0, cnt = nu, row = 0;
      bool exists = true;
      for (int b = d - 1; b >= 0; b--) {
        if (cnt <= 3) { exists = false; break; }
        int dv = cnt >> 1;
        if ((k >> b) & 1) { lo += dv; row += 2 * dv - 2; cnt -= dv; }
        else cnt = dv;
      }
      if (!exists) continue;
      Ot fl, fr;
      if (cnt <= 3) {
        leaf_case(m, sa + lo, cnt, row, fl, fr);
      } else {
        int c0 = (1 << (d + 1)) + 2 * k;
        fl = dec(nodeL[c0]);
        Ot il = dec(nodeR[c0]);
        Ot ir = dec(nodeL[c0 + 1]);
        fr = dec(nodeR[c0 + 1]);
        merge_hulls(m, fl, il, ir, fr, d & 1, row + 2 * cnt - 4, row + 2 * cnt - 3);
      }
      nodeL[nodes + k] = enc(fl);
      nodeR[nodes + k] = enc(fr);
    }
    __syncthreads();
  }
  const int rows = 2 * nu - 2;
  for (int t = tid; t < rows; t += DT)
    flag[t] = (m.getvx(t, 0) >= 0 && m.getvx(t, 1) >= 0 && m.getvx(t, 2) >= 0);
  __syncthreads();
  const int nt = block_exclusive_scan(flag, rows, s_part);
  for (int t = tid; t < rows; t += DT) {
    int v0 = m.getvx(t, 0), v1 = m.getvx(t, 1), v2 = m.getvx(t, 2);
    if (v0 >= 0 && v1 >= 0 && v2 >= 0) {
      int k = flag[t];
      tri[3 * k] = v1;       // org
      tri[3 * k + 1] = v2;   // dest
      tri[3 * k + 2] = v0;   // apex
    }
  }
  return nt;
}

// ascending bitonic sort of N (power of two) 64-bit keys in shared memory
__device__ void bitonic_sort(unsigned long long* key, int N) {
  const int tid = threadIdx.x;
  for (int k = 2; k <= N; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < N; i += DT) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = key[i], b = key[ixj];
          bool asc = (i & k) == 0;
          if ((a > b) == asc) { key[i] = b; key[ixj] = a; }
        }
      }
      __syncthreads();
    }
}

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr int OCC_EMPTY = 0x7f7f7f7f;   // cudaMemset(0x7f) pattern

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
  int* const xl_g = ws.xlist[side] + fo;
  int* const flag_g = ws.tmpD[side] + (size_t)frame * g.cap_t;  // cap_t ints
  int *xl = xl_g, *yl = ws.ylist[side] + fo, *sp = ws.tmpA[side] + fo;
  int *seglo = ws.tmpB[side] + fo, *segn = ws.tmpC[side] + fo;
  int* flag = flag_g;
  int* scan = ws.nodeL[side] + (size_t)frame * g.cap_t;         // reused before the merge phase
  const bool small = n <= g.dl_smem_max;
  if (small) {
    // ordering + partition arrays in shared memory (their passes are latency bound in HBM/L2):
    //   [0,64K) sort keys, later flag + seglo | [64K,156K) xl, yl, sp | [156K,186K) scan | [186K,216K) segn
    int* base = reinterpret_cast<int*>(dsm);
    flag = base;
    seglo = base + SMEM_MAX_POINTS;
    xl = base + 16384;
    yl = xl + SMEM_MAX_POINTS;
    sp = yl + SMEM_MAX_POINTS;
    scan = sp + SMEM_MAX_POINTS;
    segn = scan + SMEM_MAX_POINTS;
  }
  if (tid == 0) info->dt[side][0] = gtime();

  // ---- 1. (x,y) order without duplicates, then (y,x) order ------------------------------------
  int nu;
  if (n <= g.dl_sort_max) {
    unsigned long long* key = reinterpret_cast<unsigned long long*>(dsm);
    int N = 2;
    while (N < n) N <<= 1;
    for (int i = tid; i < N; i += DT)
      key[i] = (i < n) ? (((unsigned long long)px[i] << 48) | ((unsigned long long)py[i] << 32) | (unsigned)i)
                       : ~0ull;
    __syncthreads();
    bitonic_sort(key, N);
    // first of every (x,y) run survives = lowest support index
    for (int i = tid; i < n; i += DT) scan[i] = (i == 0) || ((key[i] >> 32) != (key[i - 1] >> 32));
    __syncthreads();
    nu = block_exclusive_scan(scan, n, s_part);
    for (int i = tid; i < n; i += DT) {
      bool first = (i == 0) || ((key[i] >> 32) != (key[i - 1] >> 32));
      if (first) xl[scan[i]] = (int)(key[i] & 0xFFFFFFFFu);
    }
    __syncthreads();
    for (int i = tid; i < N; i += DT) {
      unsigned long long k = ~0ull;
      if (i < nu) {
        int id = xl[i];
        k = ((unsigned long long)py[id] << 48) | ((unsigned long long)px[id] << 32) | (unsigned)id;
      }
      key[i] = k;
    }
    __syncthreads();
    bitonic_sort(key, N);
    for (int i = tid; i < nu; i += DT) yl[i] = (int)(key[i] & 0xFFFFFFFFu);
    __syncthreads();
  } else if (g.p.add_corners) {
    // corner points are off the lattice the occupancy grid is built on: not supported together
    if (tid == 0) { info->status = JN_ERR_UNSUPPORTED; info->n_tri[side] = 0; }
    return;
  } else {
    // large point sets: ranks by prefix sums over an occupancy grid (preset to OCC_EMPTY)
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
  }
  if (nu < 2) {
    if (tid == 0) info->n_tri[side] = 0;
    return;
  }
  if (tid == 0) info->dt[side][1] = gtime();

  // ---- 2. alternating-axis median partition (alternateaxes), level by level -------------
  for (int i = tid; i < nu; i += DT) { seglo[i] = 0; segn[i] = nu; }
  __syncthreads();
  int depth = 0;   // number of levels that split something = depth of the deepest leaves
  for (;; depth++) {
    if (tid == 0) s_flag = 0;
    __syncthreads();
    int* prim = (depth & 1) ? yl : xl;
    int* sec = (depth & 1) ? xl : yl;
    for (int i = tid; i < nu; i += DT) {
      int ns = segn[i];
      int r = 0;
      if (ns >= 4) { r = (i - seglo[i]) >= (ns >> 1); s_flag = 1; }
      flag[prim[i]] = r;
    }
    __syncthreads();
    if (!s_flag) break;
    for (int i = tid; i < nu; i += DT) scan[i] = flag[sec[i]];
    __syncthreads();
    block_exclusive_scan(scan, nu, s_part);
    for (int i = tid; i < nu; i += DT) {
      int ns = segn[i], lo = seglo[i], id = sec[i];
      int pos = i;
      if (ns >= 4) {
        int rb = scan[i] - scan[lo];
        pos = flag[id] ? lo + (ns >> 1) + rb : lo + (i - lo - rb);
      }
      sp[pos] = id;
    }
    __syncthreads();
    for (int i = tid; i < nu; i += DT) {
      int ns = segn[i], lo = seglo[i];
      if (ns >= 4) {
        int dv = ns >> 1;
        if (i - lo < dv) segn[i] = dv;
        else { seglo[i] = lo + dv; segn[i] = ns - dv; }
      }
    }
    // the partitioned copy becomes the secondary list
    if (depth & 1) { int* t = xl; xl = sp; sp = t; } else { int* t = yl; yl = sp; sp = t; }
    __syncthreads();
  }
  if (small) {
    // the shared memory is recycled for the triangle table: keep the final order in HBM
    for (int i = tid; i < nu; i += DT) xl_g[i] = xl[i];
    xl = xl_g;
  }
  const int* sa = xl;   // Triangle's final sortarray
  if (tid == 0) { info->dt[side][2] = gtime(); info->dmerge_depth = depth; }

  // ---- 3. merges + emission --------------------------------------------------------------------
  int* nodeL = ws.nodeL[side] + (size_t)frame * g.cap_t;
  int* nodeR = ws.nodeR[side] + (size_t)frame * g.cap_t;
  int* tri = ws.tri[side] + (size_t)frame * g.cap_t * 3;
  int nt;
  __syncthreads();
  if (n <= g.dl_smem_max) {
    MeshS m;
    m.rows = reinterpret_cast<unsigned short*>(dsm);
    unsigned* xy = reinterpret_cast<unsigned*>(dsm + (size_t)(2 * SMEM_MAX_POINTS) * 12);
    for (int i = tid; i < n; i += DT) xy[i] = ((unsigned)px[i] << 16) | (unsigned)py[i];
    m.xy = xy;
    __syncthreads();
    nt = build_and_emit(m, sa, nu, depth, nodeL, nodeR, flag_g, tri, s_part);
  } else {
    MeshG m;
    m.nb = ws.nb[side] + (size_t)frame * g.cap_t * 3;
    m.vx = ws.vx[side] + (size_t)frame * g.cap_t * 3;
    m.x = px; m.y = py;
    nt = build_and_emit(m, sa, nu, depth, nodeL, nodeR, flag_g, tri, s_part);
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
  // the occupancy grids (only used above SORT_MAX points) start "empty" = 0x7f7f7f7f
  if (g.cap_s > g.dl_sort_max)
    JN_CUDA_CHECK(cudaMemsetAsync(ws.occ, 0x7f, (size_t)B * 2 * g.W * g.Hc * sizeof(int32_t), s));
  delaunay_kernel<<<dim3(2, B), DT, DELAUNAY_SMEM, s>>>(g, ws);
  g_jn_launches += 1;
  return JN_OK;
}
