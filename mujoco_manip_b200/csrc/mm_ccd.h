// General convex-convex penetration for geom pairs that involve a collision mesh (convex hull) or a
// cylinder: GJK overlap test followed by EPA, one contact per pair (MuJoCo `multiccd` off).
// The iteration rules (start direction, simplex updates, EPA tolerances, caps and the ORDER in which
// polytope faces are removed and created) are the ones the CPU oracle states in oracle/ccd.h, so that
// both sides pick the same witness points on flat contacts.
//
// The G lanes of an env work on ONE pair at a time: hull support scans, the closest-face search, the
// visibility test and the creation of new faces are spread over the lanes; the simplex logic is replicated.
// Reductions break ties like a linear scan (lowest index), so the result does not depend on G.
#pragma once
#include "mm_model.h"

#if defined(MM_DEBUG_EPA) && !defined(__CUDA_ARCH__)
static long mm_debug_epa_iters = 0;
#endif
namespace mm {

constexpr int EPA_MAXV = 136, EPA_MAXF = 256, EPA_MAXE = 128, EPA_MAXIT = 128;
constexpr int EPA_REALS = EPA_MAXV * 6 + EPA_MAXF * 4;
constexpr int EPA_INTS = EPA_MAXF + 2 * EPA_MAXE + EPA_MAXV;
constexpr int EPA_VIS = 1 << 30;
constexpr int SHAPE_LV = 5;  // ceil(152 / 32): the largest hull has 152 vertices
// MM_HULL_REGS=1: a warp keeps its slice of both hulls in registers for the whole GJK / EPA run (support queries touch
// no memory; 60 registers); 0: support queries read the vertices from global memory (L1-resident model table)
#ifndef MM_HULL_REGS
#define MM_HULL_REGS 1
#endif
// MM_EPA_REGS=1: on the device a 32-lane group keeps the EPA visibility flags and the horizon edge list in registers
// (one edge per lane) while the horizon has at most 32 edges; the shared-memory lists are the general path
#ifndef MM_SUPPORT_TREE
#define MM_SUPPORT_TREE 1
#endif
#ifndef MM_EPA_REGS
#define MM_EPA_REGS 1
#endif
#if defined(__CUDA_ARCH__) && MM_EPA_REGS
#define EPA_REG_HORIZON(G) ((G) == 32)
#else
#define EPA_REG_HORIZON(G) false
#endif

// EPA tolerances: the oracle's values in FP64; scaled to the arithmetic's resolution in FP32 (otherwise the
// expansion never sees its progress fall below the threshold and runs into the face cap)
template <class T> MM_HD double epa_tol() { return sizeof(T) == 8 ? 1e-10 : 2e-6; }
template <class T> MM_HD double epa_vis() { return sizeof(T) == 8 ? 1e-14 : 1e-8; }

template <class T>
struct Shape {
  int type;       // GT_CYL, GT_BOX, GT_HULL
  T pos[3];       // world position
  const T* R;     // world orientation (row-major 3x3)
  T size[3];      // box half sizes / cylinder (radius, half height)
  const T* verts; // hull vertices in the geom frame (3 * nvert), global memory
  int nvert;
  // G == 32: this lane's slice of the hull (vertices lane, lane + 32, ...) held in registers for the whole
  // GJK / EPA run on the pair, so that a support query touches no memory
  T lv[MM_HULL_REGS ? SHAPE_LV : 1][3];
};

template <class T, int G>
MM_HDN void support1(const Grp<G>& g, const Shape<T>& s, const T* d, T* out) {
  const T* R = s.R;
  T l[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
            R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
  T p[3];
  if (s.type == GT_BOX) {
    p[0] = l[0] >= 0 ? s.size[0] : -s.size[0];
    p[1] = l[1] >= 0 ? s.size[1] : -s.size[1];
    p[2] = l[2] >= 0 ? s.size[2] : -s.size[2];
  } else if (s.type == GT_CYL) {
    T n = tsqrt(l[0] * l[0] + l[1] * l[1]);
    if (n > (T)1e-14) { p[0] = l[0] / n * s.size[0]; p[1] = l[1] / n * s.size[0]; } else { p[0] = p[1] = 0; }
    p[2] = l[2] >= 0 ? s.size[1] : -s.size[1];
  } else if (G == 32 && MM_HULL_REGS) {  // every hull of the model has at most 32 * SHAPE_LV vertices (static_assert in mm_model.h)
    // this lane's best of its (up to) five vertices: the first maximum, found by a tournament (pairs, then pairs of
    // winners: three dependent compare / select levels instead of five; a later vertex wins only when strictly greater,
    // which is the tie rule of the linear scan); lanes without a vertex lose every comparison
    static_assert(SHAPE_LV == 5, "tournament below is written for five slices");
    T v[SHAPE_LV];
    int ix[SHAPE_LV];
#pragma unroll
    for (int k = 0; k < SHAPE_LV; k++) {
      ix[k] = g.lane + 32 * k;
      v[k] = s.lv[k][0] * l[0] + s.lv[k][1] * l[1] + s.lv[k][2] * l[2];
      if (ix[k] >= s.nvert) { v[k] = (T)-1e30; ix[k] = s.nvert; }
    }
#if MM_SUPPORT_TREE
    { const bool t = v[1] > v[0]; v[0] = t ? v[1] : v[0]; ix[0] = t ? ix[1] : ix[0]; }
    { const bool t = v[3] > v[2]; v[2] = t ? v[3] : v[2]; ix[2] = t ? ix[3] : ix[2]; }
    { const bool t = v[2] > v[0]; v[0] = t ? v[2] : v[0]; ix[0] = t ? ix[2] : ix[0]; }
    { const bool t = v[4] > v[0]; v[0] = t ? v[4] : v[0]; ix[0] = t ? ix[4] : ix[0]; }
#else
#pragma unroll
    for (int k = 1; k < SHAPE_LV; k++) { const bool t = v[k] > v[0]; v[0] = t ? v[k] : v[0]; ix[0] = t ? ix[k] : ix[0]; }
#endif
    T bv = v[0];
    int best = ix[0];
    g.argmax_index(bv, best);
    if (best >= s.nvert) best = 0;
    int kb = best >> 5;
    T q[3] = {s.lv[0][0], s.lv[0][1], s.lv[0][2]};
#pragma unroll
    for (int k = 1; k < SHAPE_LV; k++)
      if (k == kb) { q[0] = s.lv[k][0]; q[1] = s.lv[k][1]; q[2] = s.lv[k][2]; }
    p[0] = g.bcast(q[0], best & 31); p[1] = g.bcast(q[1], best & 31); p[2] = g.bcast(q[2], best & 31);
  } else {
    int best = s.nvert;  // lanes without a vertex lose every comparison
    T bv = (T)-1e30;
    const T* V = s.verts;
    for (int i = g.lane; i < s.nvert; i += G) {
      T v = V[3 * i] * l[0] + V[3 * i + 1] * l[1] + V[3 * i + 2] * l[2];
      if (v > bv) { bv = v; best = i; }
    }
    g.argmax_index(bv, best);
    if (best >= s.nvert) best = 0;
    p[0] = V[3 * best]; p[1] = V[3 * best + 1]; p[2] = V[3 * best + 2];
  }
  out[0] = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + s.pos[0];
  out[1] = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + s.pos[1];
  out[2] = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + s.pos[2];
}

// support point of the Minkowski difference: v = a - b
template <class T>
struct SP { T v[3], a[3]; };

template <class T, int G>
MM_HDN void support(const Grp<G>& g, const Shape<T>& s1, const Shape<T>& s2, const T* d, SP<T>& p) {
  T nd[3] = {-d[0], -d[1], -d[2]}, b[3];
  support1<T, G>(g, s1, d, p.a);
  support1<T, G>(g, s2, nd, b);
  for (int k = 0; k < 3; k++) p.v[k] = p.a[k] - b[k];
}

template <class T> MM_HD void sub3(T* r, const T* a, const T* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <class T> MM_HD void neg3(T* r, const T* a) { r[0] = -a[0]; r[1] = -a[1]; r[2] = -a[2]; }
template <class T> MM_HD void cpy3(T* r, const T* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }

// true if the shapes overlap; sx[0..3] then hold a tetrahedron around the origin (same on every lane)
template <class T, int G>
MM_HDN bool gjk(const Grp<G>& g, const Shape<T>& s1, const Shape<T>& s2, SP<T>* sx, unsigned* cnt = nullptr) {
  T dir[3] = {s2.pos[0] - s1.pos[0], s2.pos[1] - s1.pos[1], s2.pos[2] - s1.pos[2]};
  if (dot3(dir, dir) < (T)1e-20) { dir[0] = 1; dir[1] = 0; dir[2] = 0; }
  // ONE support call site for the whole loop (the two opening queries are its first two turns): the support scan is the
  // bulk of this kernel's code, and every inlined copy costs instruction cache
  SP<T> a, b, c, d;
  T bc[3], nb[3], t[3];
  int n = 0;  // 0, 1: opening queries; 2: line / triangle case; 3: tetrahedron case
#pragma unroll 1
  for (int it = -2; it < 64; it++) {
    if (cnt && g.lane == 0 && it >= 0) cnt[1]++;
    support<T, G>(g, s1, s2, dir, a);
    if (dot3(a.v, dir) < 0) return false;
    if (n == 0) {
      c = a;
      neg3(dir, c.v);
      if (dot3(dir, dir) < (T)1e-24) { dir[0] = 1; dir[1] = 0; dir[2] = 0; }
      n = 1;
      continue;
    }
    if (n == 1) {
      b = a;
      sub3(bc, c.v, b.v);
      neg3(nb, b.v);
      cross3(t, bc, nb);
      cross3(dir, t, bc);
      if (dot3(dir, dir) < (T)1e-24) {
        const T ex[3] = {1, 0, 0}, ez[3] = {0, 0, 1};
        cross3(dir, bc, ex);
        if (dot3(dir, dir) < (T)1e-24) cross3(dir, bc, ez);
      }
      n = 2;
      d = c;
      continue;
    }
    T ao[3], ab[3], ac[3];
    neg3(ao, a.v);
    sub3(ab, b.v, a.v);
    sub3(ac, c.v, a.v);
    if (n == 2) {
      T nrm[3], e[3];
      cross3(nrm, ab, ac);
      cross3(e, ab, nrm);
      if (dot3(e, ao) > 0) {
        c = a;
        cross3(t, ab, ao);
        cross3(dir, t, ab);
        if (dot3(dir, dir) < (T)1e-24) cpy3(dir, nrm);
        continue;
      }
      cross3(e, nrm, ac);
      if (dot3(e, ao) > 0) {
        b = a;
        cross3(t, ac, ao);
        cross3(dir, t, ac);
        if (dot3(dir, dir) < (T)1e-24) cpy3(dir, nrm);
        continue;
      }
      n = 3;
      if (dot3(nrm, ao) > 0) { d = c; c = b; b = a; cpy3(dir, nrm); }
      else { d = b; b = a; neg3(dir, nrm); }
      continue;
    }
    T ad[3], abc[3], acd[3], adb[3];
    sub3(ad, d.v, a.v);
    cross3(abc, ab, ac);
    cross3(acd, ac, ad);
    cross3(adb, ad, ab);
    if (dot3(abc, ao) > 0) { d = c; c = b; b = a; cpy3(dir, abc); continue; }
    if (dot3(acd, ao) > 0) { b = a; cpy3(dir, acd); continue; }
    if (dot3(adb, ao) > 0) { c = d; d = b; b = a; cpy3(dir, adb); continue; }
    sx[0] = a; sx[1] = b; sx[2] = c; sx[3] = d;
    return true;
  }
  return false;
}

// EPA polytope in the env's workspace (global memory, shared by the lanes of the group)
template <class T>
struct EpaMem {
  T* vert;    // [EPA_MAXV][6]  v(3), a(3)
  T* face;    // [EPA_MAXF][4]  n(3), d
  int* fidx;  // [EPA_MAXF]     three vertex indices packed 10 bits each (+ EPA_VIS while a face is being removed)
  int* edge;  // [EPA_MAXE]     horizon edges: two vertex indices packed 16 bits each
  int* canon; // [EPA_MAXV]     lowest vertex index with identical coordinates (edge matching compares coordinates)
  int* ecan;  // [EPA_MAXE]     canonical ids of the horizon edges' end points, packed like `edge`
};

// face f from vertices (ia, ib, ic); flip = which two indices swap when the normal points inward
template <class T>
MM_HD void epa_mkface(const EpaMem<T>& m, int f, int ia, int ib, int ic, bool initial) {
  const T *a = m.vert + 6 * ia, *b = m.vert + 6 * ib, *c = m.vert + 6 * ic;
  T e1[3], e2[3], n[3];
  sub3(e1, b, a);
  sub3(e2, c, a);
  cross3(n, e1, e2);
  T l = tsqrt(dot3(n, n));
  T* F = m.face + 4 * f;
  if ((double)l < 1e-30) { F[0] = F[1] = F[2] = 0; F[3] = (T)1e30; m.fidx[f] = ia | (ib << 10) | (ic << 20); return; }
  T inv = (T)1 / l;
  n[0] *= inv; n[1] *= inv; n[2] *= inv;
  T d = dot3(n, a);
  if (d < 0) {  // wind outward: the oracle swaps p[1], p[2] on the initial tetrahedron and p[0], p[1] on later faces
    F[0] = -n[0]; F[1] = -n[1]; F[2] = -n[2]; F[3] = -d;
    m.fidx[f] = initial ? (ia | (ic << 10) | (ib << 20)) : (ib | (ia << 10) | (ic << 20));
  } else {
    F[0] = n[0]; F[1] = n[1]; F[2] = n[2]; F[3] = d;
    m.fidx[f] = ia | (ib << 10) | (ic << 20);
  }
}

// closest face: minimal d, lowest index among equals
template <class T, int G>
MM_HD int epa_best(const Grp<G>& g, const EpaMem<T>& m, int nf) {
  T bv = (T)3e38;
  int best = nf;
#pragma unroll 2
  for (int i = g.lane; i < nf; i += G) {
    const T d = m.face[4 * i + 3];
    const bool lt = d < bv;  // (select form: no data-dependent branch in the scan)
    bv = lt ? d : bv;
    best = lt ? i : best;
  }
  g.argmin_index(bv, best);
  return best >= nf ? 0 : best;
}

// Outputs contact position (mid witness), normal (shape1 -> shape2) and penetration depth (same on every lane).
template <class T, int G>
MM_HDN bool epa(const Grp<G>& g, const Shape<T>& s1, const Shape<T>& s2, const SP<T>* sx, const EpaMem<T>& m, T* pos,
                T* nrm, T* depth, unsigned* cnt = nullptr) {
  int nv = 4, nf = 4;
  if (g.lane == 0) {
    for (int i = 0; i < 4; i++) {
      for (int k = 0; k < 3; k++) { m.vert[6 * i + k] = sx[i].v[k]; m.vert[6 * i + 3 + k] = sx[i].a[k]; }
      int cn = i;
      for (int j = i - 1; j >= 0; j--)
        if (sx[j].v[0] == sx[i].v[0] && sx[j].v[1] == sx[i].v[1] && sx[j].v[2] == sx[i].v[2]) cn = j;
      m.canon[i] = cn;
    }
  }
  g.sync();
  // the four faces of the tetrahedron (0,1,2), (0,2,3), (0,3,1), (1,3,2) on four lanes (each is a square root and a
  // division deep: one after the other on a single lane they were ~1,000 cycles of every EPA run)
  for (int f = g.lane; f < 4; f += G)
    epa_mkface(m, f, f == 3 ? 1 : 0, f == 0 ? 1 : (f == 1 ? 2 : 3), f == 0 ? 2 : (f == 1 ? 3 : (f == 2 ? 1 : 2)), true);
  g.sync();
#pragma unroll 1
  for (int it = 0; it < EPA_MAXIT; it++) {
    if (cnt && g.lane == 0) cnt[3]++;
#if defined(MM_DEBUG_EPA) && !defined(__CUDA_ARCH__)
    mm_debug_epa_iters++;
#endif
    int best = epa_best<T, G>(g, m, nf);
    T n[3] = {m.face[4 * best], m.face[4 * best + 1], m.face[4 * best + 2]};
    SP<T> p;
    support<T, G>(g, s1, s2, n, p);
    T dist = dot3(p.v, n);
    if ((double)(dist - m.face[4 * best + 3]) < epa_tol<T>() || nf >= EPA_MAXF - 8) break;
    int ip = nv++;
    // new vertex, its canonical index, and the visibility flag of every face (spread over the lanes)
    int cn = ip;
    for (int j = g.lane; j < ip; j += G) {
      const T* q = m.vert + 6 * j;
      const bool same = q[0] == p.v[0] && q[1] == p.v[1] && q[2] == p.v[2] && j < cn;
      cn = same ? j : cn;
    }
    cn = g.imin(cn);
    if (g.lane == 0) {
      for (int k = 0; k < 3; k++) { m.vert[6 * ip + k] = p.v[k]; m.vert[6 * ip + 3 + k] = p.a[k]; }
      m.canon[ip] = cn;
    }
    // (32 lanes on the device: the visibility flags stay in a register - bit j of lane l = face l + 32 j - and go to
    // the faces' index words only if the horizon cannot be held in registers, see below)
    unsigned visbits = 0;
    for (int i = g.lane, j = 0; i < nf; i += G, j++) {
      int fi = m.fidx[i];
      const T* F = m.face + 4 * i;
      T r[3];
      sub3(r, p.v, m.vert + 6 * (fi & 1023));
      if ((double)dot3(F, r) > epa_vis<T>()) {
        if (EPA_REG_HORIZON(G)) visbits |= 1u << j; else m.fidx[i] = fi | EPA_VIS;
      }
    }
    // Removal of the visible faces and collection of the horizon.  The SEQUENCE of removals (face i is replaced by the
    // last face and examined again) and of edge insertions / cancellations is the oracle's - it fixes the order of the
    // new faces and therefore every later tie - but each step is done by all lanes: the next visible face and the
    // matching reversed edge are found with a ballot instead of a scan by one lane.
    int ne = 0, ereg = 0;
    bool inregs = false;
#if defined(__CUDA_ARCH__) && MM_EPA_REGS
    if (EPA_REG_HORIZON(G)) {
      // Horizon in registers: at most 3 edges per visible face, so with <= 10 visible faces the edge list fits one entry
      // per lane (lane k = edge k: end points and their canonical ids).  Same sequence as the general path (lowest
      // matching entry by ballot + count-trailing-zeros, the last entry moves into the hole by a shuffle), but an edge
      // step touches no shared memory and needs no barrier; the only barrier left is the one behind a face's move into
      // the hole of a removed one.  More visible faces: flags to the index words, general path.
      inregs = 3 * g.isum(__popc(visbits)) <= 32;
      if (!inregs)
        for (int i = g.lane, j = 0; i < nf; i += G, j++)
          if ((visbits >> j) & 1u) m.fidx[i] |= EPA_VIS;
    }
#endif
    g.sync();
    if (inregs) {
#if defined(__CUDA_ARCH__) && MM_EPA_REGS
      int creg = -1;
#pragma unroll 1
      for (int i = 0;;) {
        int found_face = -1;
#pragma unroll 1
        for (int j = i >> 5; (j << 5) < nf; j++) {
          const int idx = (j << 5) + g.lane;
          const unsigned b = g.ballot(idx >= i && idx < nf && ((visbits >> j) & 1u));
          if (b) { found_face = (j << 5) + tctz(b); break; }
        }
        if (found_face < 0) break;
        i = found_face;
        const int fi = m.fidx[i];
        const int id[3] = {fi & 1023, (fi >> 10) & 1023, (fi >> 20) & 1023};
        const int cid[3] = {m.canon[id[0]], m.canon[id[1]], m.canon[id[2]]};
#pragma unroll
        for (int e = 0; e < 3; e++) {
          const int f = e == 2 ? 0 : e + 1;
          const int rev = cid[f] | (cid[e] << 16);  // shared edges appear reversed
          const unsigned b = g.ballot(g.lane < ne && creg == rev);
          // (select form: both outcomes cost a ballot and two shuffles, no branch)
          const int hit = b ? tctz(b) : -1;
          ne -= b ? 1 : 0;
          const int le = g.bcast(ereg, ne & 31), lc = g.bcast(creg, ne & 31);
          const bool take = g.lane == hit, put = !b && g.lane == ne;
          ereg = take ? le : (put ? (id[e] | (id[f] << 16)) : ereg);
          creg = take ? lc : (put ? (cid[e] | (cid[f] << 16)) : creg);
          ne += b ? 0 : 1;
        }
        --nf;
        for (int k = g.lane; k < 4; k += G) m.face[4 * i + k] = m.face[4 * nf + k];
        if (g.lane == G - 1) m.fidx[i] = m.fidx[nf];
        const unsigned vmoved = (g.bcast(visbits, nf & 31) >> (nf >> 5)) & 1u;
        if (g.lane == (i & 31)) visbits = (visbits & ~(1u << (i >> 5))) | (vmoved << (i >> 5));
        if (g.lane == (nf & 31)) visbits &= ~(1u << (nf >> 5));
        g.sync();
      }
#endif
    } else {
      for (int i = 0;;) {
        // next visible face at or after i
        int found_face = -1;
        for (int base = i; base < nf; base += G) {
          int idx = base + g.lane;
          unsigned b = g.ballot(idx < nf && (m.fidx[idx] & EPA_VIS));
          if (b) { found_face = base + tctz(b); break; }
        }
        if (found_face < 0) break;
        i = found_face;
        int fi = m.fidx[i];
        int id[3] = {fi & 1023, (fi >> 10) & 1023, (fi >> 20) & 1023};
#pragma unroll 1
        for (int e = 0; e < 3; e++) {
          int ea = id[e], eb = id[e == 2 ? 0 : e + 1];
          int rev = m.canon[eb] | (m.canon[ea] << 16);  // shared edges appear reversed
          int hit = -1;
          for (int base = 0; base < ne; base += G) {
            int k = base + g.lane;
            unsigned b = g.ballot(k < ne && m.ecan[k] == rev);
            if (b) { hit = base + tctz(b); break; }
          }
          if (hit >= 0) {
            --ne;
            if (g.lane == 0) { m.edge[hit] = m.edge[ne]; m.ecan[hit] = m.ecan[ne]; }
          } else if (ne < EPA_MAXE) {
            if (g.lane == 0) { m.edge[ne] = ea | (eb << 16); m.ecan[ne] = m.canon[ea] | (m.canon[eb] << 16); }
            ne++;
          }
          g.sync();
        }
        --nf;
        for (int k = g.lane; k < 4; k += G) m.face[4 * i + k] = m.face[4 * nf + k];
        if (g.lane == G - 1) m.fidx[i] = m.fidx[nf];
        g.sync();
      }
    }
    if (ne == 0) break;
    int add = ne < EPA_MAXF - nf ? ne : EPA_MAXF - nf;
    for (int k = g.lane; k < add; k += G) {
      const int ed = inregs ? ereg : m.edge[k];  // (register list: at most 32 edges, lane k holds edge k)
      epa_mkface(m, nf + k, ed & 0xFFFF, ed >> 16, ip, false);
    }
    nf += add;
    g.sync();
  }
  int best = epa_best<T, G>(g, m, nf);
  const T* F = m.face + 4 * best;
  if (!((double)F[3] < 1e29)) return false;
  int fi = m.fidx[best];
  const T *p0 = m.vert + 6 * (fi & 1023), *p1 = m.vert + 6 * ((fi >> 10) & 1023), *p2 = m.vert + 6 * ((fi >> 20) & 1023);
  T pr[3] = {F[0] * F[3], F[1] * F[3], F[2] * F[3]}, v0[3], v1[3], v2[3];
  sub3(v0, p1, p0);
  sub3(v1, p2, p0);
  sub3(v2, pr, p0);
  T d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
  T den = d00 * d11 - d01 * d01, v = 0, w = 0;
  if ((double)tabs(den) > 1e-30) { v = (d11 * d20 - d01 * d21) / den; w = (d00 * d21 - d01 * d20) / den; }
  T u = 1 - v - w;
  for (int k = 0; k < 3; k++) {
    T wa = p0[3 + k] * u + p1[3 + k] * v + p2[3 + k] * w;
    T wb = (p0[3 + k] - p0[k]) * u + (p1[3 + k] - p1[k]) * v + (p2[3 + k] - p2[k]) * w;
    pos[k] = (T)0.5 * (wa + wb);
    nrm[k] = F[k];
  }
  *depth = F[3];
  return F[3] > 0;
}

}  // namespace mm
