/*
 * delaunay_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the one path of Shewchuk's Triangle 1.6 that the reference
 * executes: triangulate("zQB") = divide-and-conquer Delaunay with alternating
 * (Dwyer) cuts on a triangle-based structure with ghost triangles, exact
 * predicates, no jettisoning (vertex numbers = input order).
 *
 * The third-party algorithm is vendored in the reference as
 * src/elas/triangle.cpp (Triangle 1.6, REAL = float); lines restated:
 *   vertexsort 5446-5499, vertexmedian 5513-5569, alternateaxes 5582-5601,
 *   mergehulls 5638-5934, divconqrecurse 5953-6103, removeghosts 6105-6148,
 *   divconqdelaunay 6160-6217, randomnation 4045-4049 (seed reset 4030),
 *   predicates counterclockwise 2706-2744 / incircle 3334-3381,
 *   pool order poolalloc 1709-1757 / traverse 1812-1842, writeelements 7800-7862.
 *
 * Differences in mechanism, not in result:
 *   - triangles are rows of two int tables (3 neighbour handles = 4*index +
 *     orientation, 3 vertex ids, -1 = the ghost vertex) instead of pointer
 *     blocks; the row index equals Triangle's allocation order, so emitting
 *     the non-ghost rows in index order reproduces writeelements' order and
 *     corner rotation;
 *   - predicates are evaluated exactly in integers.  The caller's coordinates
 *     are integer-valued floats (pixel positions), for which Triangle's
 *     adaptive float arithmetic returns the same sign.
 * Pinned against the real Triangle by tests/test_oracle_pin.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int t, o; } Ot; /* oriented triangle: row + orientation 0..2 */

typedef struct {
  int* nb;     /* [3*T] neighbour handles */
  int* vx;     /* [3*T] vertex ids        */
  int  ntri;   /* rows allocated so far   */
  const int32_t* x;
  const int32_t* y;
  uint64_t seed;
} Mesh;

static const int P1[3] = {1, 2, 0};
static const int M1[3] = {2, 0, 1};

static inline int enc(Ot a) { return a.t * 4 + a.o; }
static inline Ot dec(int e) { Ot r; r.t = e >> 2; r.o = e & 3; return r; }
static inline Ot sym(const Mesh* m, Ot a) { return dec(m->nb[3 * a.t + a.o]); }
static inline Ot lnext(Ot a) { a.o = P1[a.o]; return a; }
static inline Ot lprev(Ot a) { a.o = M1[a.o]; return a; }
static inline int org(const Mesh* m, Ot a) { return m->vx[3 * a.t + P1[a.o]]; }
static inline int dest(const Mesh* m, Ot a) { return m->vx[3 * a.t + M1[a.o]]; }
static inline int apex(const Mesh* m, Ot a) { return m->vx[3 * a.t + a.o]; }
static inline void setorg(Mesh* m, Ot a, int v) { m->vx[3 * a.t + P1[a.o]] = v; }
static inline void setdest(Mesh* m, Ot a, int v) { m->vx[3 * a.t + M1[a.o]] = v; }
static inline void setapex(Mesh* m, Ot a, int v) { m->vx[3 * a.t + a.o] = v; }
static inline void bond(Mesh* m, Ot a, Ot b) {
  m->nb[3 * a.t + a.o] = enc(b);
  m->nb[3 * b.t + b.o] = enc(a);
}

static Ot maketri(Mesh* m) {
  Ot r;
  r.t = m->ntri++;
  r.o = 0;
  for (int k = 0; k < 3; k++) { m->nb[3 * r.t + k] = -1; m->vx[3 * r.t + k] = -1; }
  return r;
}

/* sign of the orientation determinant; > 0 if a,b,c counterclockwise */
static inline int64_t ccw(const Mesh* m, int a, int b, int c) {
  int64_t ax = m->x[a] - m->x[c], ay = m->y[a] - m->y[c];
  int64_t bx = m->x[b] - m->x[c], by = m->y[b] - m->y[c];
  return ax * by - ay * bx;
}

/* > 0 if d lies inside the circle through a,b,c (counterclockwise) */
static inline int incircle_pos(const Mesh* m, int a, int b, int c, int d) {
  __int128 adx = m->x[a] - m->x[d], ady = m->y[a] - m->y[d];
  __int128 bdx = m->x[b] - m->x[d], bdy = m->y[b] - m->y[d];
  __int128 cdx = m->x[c] - m->x[d], cdy = m->y[c] - m->y[d];
  __int128 al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
  __int128 det = al * (bdx * cdy - cdx * bdy) + bl * (cdx * ady - adx * cdy) + cl * (adx * bdy - bdx * ady);
  return det > 0;
}

static unsigned long long randomnation(Mesh* m, unsigned int choices) {
  m->seed = (m->seed * 1366ull + 150889ull) % 714025ull;
  return m->seed / (714025ull / choices + 1);
}

/* lexicographic compare on (axis, other axis) */
static inline int lt(const Mesh* m, int a, int32_t k1, int32_t k2, int axis) {
  int32_t a1 = axis ? m->y[a] : m->x[a], a2 = axis ? m->x[a] : m->y[a];
  return a1 < k1 || (a1 == k1 && a2 < k2);
}
static inline int gt(const Mesh* m, int a, int32_t k1, int32_t k2, int axis) {
  int32_t a1 = axis ? m->y[a] : m->x[a], a2 = axis ? m->x[a] : m->y[a];
  return a1 > k1 || (a1 == k1 && a2 > k2);
}

/* randomized quicksort by (x,y); the pivot sequence decides which of two
 * duplicate points comes first, hence it is reproduced exactly */
static void vertexsort(Mesh* m, int* a, int n) {
  if (n == 2) {
    if (gt(m, a[0], m->x[a[1]], m->y[a[1]], 0)) { int t = a[1]; a[1] = a[0]; a[0] = t; }
    return;
  }
  int pivot = (int)randomnation(m, (unsigned)n);
  int32_t px = m->x[a[pivot]], py = m->y[a[pivot]];
  int left = -1, right = n;
  while (left < right) {
    do { left++; } while (left <= right && lt(m, a[left], px, py, 0));
    do { right--; } while (left <= right && gt(m, a[right], px, py, 0));
    if (left < right) { int t = a[left]; a[left] = a[right]; a[right] = t; }
  }
  if (left > 1) vertexsort(m, a, left);
  if (right < n - 2) vertexsort(m, a + right + 1, n - right - 1);
}

static void vertexmedian(Mesh* m, int* a, int n, int median, int axis) {
  if (n == 2) {
    int32_t k1 = axis ? m->y[a[1]] : m->x[a[1]], k2 = axis ? m->x[a[1]] : m->y[a[1]];
    if (gt(m, a[0], k1, k2, axis)) { int t = a[1]; a[1] = a[0]; a[0] = t; }
    return;
  }
  int pivot = (int)randomnation(m, (unsigned)n);
  int32_t k1 = axis ? m->y[a[pivot]] : m->x[a[pivot]], k2 = axis ? m->x[a[pivot]] : m->y[a[pivot]];
  int left = -1, right = n;
  while (left < right) {
    do { left++; } while (left <= right && lt(m, a[left], k1, k2, axis));
    do { right--; } while (left <= right && gt(m, a[right], k1, k2, axis));
    if (left < right) { int t = a[left]; a[left] = a[right]; a[right] = t; }
  }
  if (left > median) vertexmedian(m, a, left, median, axis);
  if (right < median - 1) vertexmedian(m, a + right + 1, n - right - 1, median - right - 1, axis);
}

static void alternateaxes(Mesh* m, int* a, int n, int axis) {
  int divider = n >> 1;
  if (n <= 3) axis = 0;
  vertexmedian(m, a, n, divider, axis);
  if (n - divider >= 2) {
    if (divider >= 2) alternateaxes(m, a, divider, 1 - axis);
    alternateaxes(m, a + divider, n - divider, 1 - axis);
  }
}

/* Knit two triangulations along the gap between them. */
static void mergehulls(Mesh* m, Ot* farleft, Ot* innerleft, Ot* innerright, Ot* farright, int axis) {
  const int32_t* X = m->x; const int32_t* Y = m->y;
  int ild = dest(m, *innerleft), ila = apex(m, *innerleft);
  int iro = org(m, *innerright), ira = apex(m, *innerright);
  Ot check;
  int cv;
  if (axis == 1) {
    /* horizontal cut: walk the four handles to the bottom-/topmost hull vertices */
    int flp = org(m, *farleft), fla = apex(m, *farleft);
    int frp = dest(m, *farright), fra = apex(m, *farright);
    while (Y[fla] < Y[flp]) {
      *farleft = sym(m, lnext(*farleft));
      flp = fla;
      fla = apex(m, *farleft);
    }
    check = sym(m, *innerleft);
    cv = apex(m, check);
    while (Y[cv] > Y[ild]) {
      *innerleft = lnext(check);
      ila = ild;
      ild = cv;
      check = sym(m, *innerleft);
      cv = apex(m, check);
    }
    while (Y[ira] < Y[iro]) {
      *innerright = sym(m, lnext(*innerright));
      iro = ira;
      ira = apex(m, *innerright);
    }
    check = sym(m, *farright);
    cv = apex(m, check);
    while (Y[cv] > Y[frp]) {
      *farright = lnext(check);
      fra = frp;
      frp = cv;
      check = sym(m, *farright);
      cv = apex(m, check);
    }
    (void)fra;
  }
  /* lower common tangent */
  int changed;
  do {
    changed = 0;
    if (ccw(m, ild, ila, iro) > 0) {
      *innerleft = sym(m, lprev(*innerleft));
      ild = ila;
      ila = apex(m, *innerleft);
      changed = 1;
    }
    if (ccw(m, ira, iro, ild) > 0) {
      *innerright = sym(m, lnext(*innerright));
      iro = ira;
      ira = apex(m, *innerright);
      changed = 1;
    }
  } while (changed);
  Ot leftcand = sym(m, *innerleft), rightcand = sym(m, *innerright);
  /* bottom ghost */
  Ot base = maketri(m);
  bond(m, base, *innerleft);
  base = lnext(base);
  bond(m, base, *innerright);
  base = lnext(base);
  setorg(m, base, iro);
  setdest(m, base, ild);
  if (ild == org(m, *farleft)) *farleft = lnext(base);
  if (iro == dest(m, *farright)) *farright = lprev(base);
  int ll = ild, lr = iro;
  int ul = apex(m, leftcand), ur = apex(m, rightcand);
  for (;;) {
    int leftdone = ccw(m, ul, ll, lr) <= 0;
    int rightdone = ccw(m, ur, ll, lr) <= 0;
    if (leftdone && rightdone) {
      /* top ghost */
      Ot top = maketri(m);
      setorg(m, top, ll);
      setdest(m, top, lr);
      bond(m, top, base);
      top = lnext(top);
      bond(m, top, rightcand);
      top = lnext(top);
      bond(m, top, leftcand);
      if (axis == 1) {
        /* restore leftmost / rightmost handles */
        int flp = org(m, *farleft), frp = dest(m, *farright), fra = apex(m, *farright);
        check = sym(m, *farleft);
        cv = apex(m, check);
        while (X[cv] < X[flp]) {
          *farleft = lprev(check);
          flp = cv;
          check = sym(m, *farleft);
          cv = apex(m, check);
        }
        while (X[fra] > X[frp]) {
          *farright = sym(m, lprev(*farright));
          frp = fra;
          fra = apex(m, *farright);
        }
      }
      return;
    }
    if (!leftdone) {
      Ot nxt = sym(m, lprev(leftcand));
      int na = apex(m, nxt);
      if (na != -1) {
        int bad = incircle_pos(m, ll, lr, ul, na);
        while (bad) {
          /* flip away the left edge: one more ghost on the left hull */
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
          nxt = sidec;
          na = apex(m, nxt);
          bad = (na != -1) ? incircle_pos(m, ll, lr, ul, na) : 0;
        }
      }
    }
    if (!rightdone) {
      Ot nxt = sym(m, lnext(rightcand));
      int na = apex(m, nxt);
      if (na != -1) {
        int bad = incircle_pos(m, ll, lr, ur, na);
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
          nxt = sidec;
          na = apex(m, nxt);
          bad = (na != -1) ? incircle_pos(m, ll, lr, ur, na) : 0;
        }
      }
    }
    if (leftdone || (!rightdone && incircle_pos(m, ul, ll, lr, ur))) {
      /* new edge ll -> ur */
      bond(m, base, rightcand);
      base = lprev(rightcand);
      setdest(m, base, ll);
      lr = ur;
      rightcand = sym(m, base);
      ur = apex(m, rightcand);
    } else {
      /* new edge ul -> lr */
      bond(m, base, leftcand);
      base = lnext(leftcand);
      setorg(m, base, lr);
      ll = ul;
      leftcand = sym(m, base);
      ul = apex(m, leftcand);
    }
  }
}

static void divconq(Mesh* m, const int* sa, int n, int axis, Ot* farleft, Ot* farright) {
  if (n == 2) {
    /* an edge = two ghosts */
    Ot a = maketri(m);
    setorg(m, a, sa[0]);
    setdest(m, a, sa[1]);
    Ot b = maketri(m);
    setorg(m, b, sa[1]);
    setdest(m, b, sa[0]);
    bond(m, a, b);
    a = lprev(a); b = lnext(b);
    bond(m, a, b);
    a = lprev(a); b = lnext(b);
    bond(m, a, b);
    *farright = b;
    *farleft = lprev(b);
  } else if (n == 3) {
    Ot mid = maketri(m), t1 = maketri(m), t2 = maketri(m), t3 = maketri(m);
    int64_t area = ccw(m, sa[0], sa[1], sa[2]);
    if (area == 0) {
      /* collinear: two edges, four ghosts */
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
      *farleft = t1;
      *farright = t2;
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
      *farleft = t1;
      *farright = (area > 0) ? t2 : lnext(*farleft);
    }
  } else {
    int divider = n >> 1;
    Ot innerleft, innerright;
    divconq(m, sa, divider, 1 - axis, farleft, &innerleft);
    divconq(m, sa + divider, n - divider, 1 - axis, &innerright, farright);
    mergehulls(m, farleft, &innerleft, &innerright, farright, axis);
  }
}

/* xy: n points as integer-valued floats.  Writes (c1,c2,c3) vertex-number
 * triples in Triangle's output order; returns the triangle count. */
int port_triangulate(const float* xy, int n, int32_t* tri_out, int cap_tri) {
  if (n < 3) return 0;
  Mesh m;
  int32_t* x = (int32_t*)malloc(sizeof(int32_t) * n);
  int32_t* y = (int32_t*)malloc(sizeof(int32_t) * n);
  for (int i = 0; i < n; i++) { x[i] = (int32_t)xy[2 * i]; y[i] = (int32_t)xy[2 * i + 1]; }
  m.x = x; m.y = y;
  m.seed = 1;
  m.ntri = 0;
  m.nb = (int*)malloc(sizeof(int) * 3 * (2 * (size_t)n + 4));
  m.vx = (int*)malloc(sizeof(int) * 3 * (2 * (size_t)n + 4));
  int* sa = (int*)malloc(sizeof(int) * n);
  for (int i = 0; i < n; i++) sa[i] = i;
  vertexsort(&m, sa, n);
  int k = 0; /* drop duplicates: the first of each run survives */
  for (int j = 1; j < n; j++)
    if (!(x[sa[k]] == x[sa[j]] && y[sa[k]] == y[sa[j]])) sa[++k] = sa[j];
  k++;
  int nt = 0;
  if (k >= 2) {
    int divider = k >> 1;
    if (k - divider >= 2) {
      if (divider >= 2) alternateaxes(&m, sa, divider, 1);
      alternateaxes(&m, sa + divider, k - divider, 1);
    }
    Ot hl, hr;
    divconq(&m, sa, k, 0, &hl, &hr);
    /* ghosts are the rows with a -1 vertex; the rest, in row order, with
     * orientation 0: org = vx[1], dest = vx[2], apex = vx[0] */
    for (int t = 0; t < m.ntri; t++) {
      const int* v = m.vx + 3 * t;
      if (v[0] < 0 || v[1] < 0 || v[2] < 0) continue;
      if (nt < cap_tri) { tri_out[3 * nt] = v[1]; tri_out[3 * nt + 1] = v[2]; tri_out[3 * nt + 2] = v[0]; }
      nt++;
    }
  }
  free(sa); free(m.nb); free(m.vx); free(x); free(y);
  return nt;
}
