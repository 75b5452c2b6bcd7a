// General convex-convex penetration for geom pairs that involve a collision mesh (convex hull) or a
// cylinder: GJK overlap test followed by EPA, one contact per pair (MuJoCo `multiccd` off).
// The iteration rules (start direction, simplex updates, EPA tolerances and caps) are the ones the CPU
// oracle states in oracle/ccd.h, so that both sides pick the same witness points on flat contacts.
//
// A lane runs GJK for its own pair (registers only).  EPA needs a polytope (<=132 vertices, <=256 faces);
// it lives in the env's global workspace and pairs that need it are expanded one at a time.
#pragma once
#include "mm_model.h"

namespace mm {

constexpr int EPA_MAXV = 136, EPA_MAXF = 256, EPA_MAXE = 128, EPA_MAXIT = 128;
constexpr int EPA_REALS = EPA_MAXV * 6 + EPA_MAXF * 4;
constexpr int EPA_INTS = EPA_MAXF + EPA_MAXE;

template <class T>
struct Shape {
  int type;       // GT_CYL, GT_BOX, GT_HULL
  T pos[3];       // world position
  const T* R;     // world orientation (row-major 3x3)
  T size[3];      // box half sizes / cylinder (radius, half height)
  const T* verts; // hull vertices in the geom frame (3 * nvert), global memory
  int nvert;
};

template <class T>
MM_HDX void support1(const Shape<T>& s, const T* d, T* out) {
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
  } else {
    int best = 0;
    T bv = (T)-1e30;
    const T* V = s.verts;
    for (int i = 0; i < s.nvert; i++) {
      T v = V[3 * i] * l[0] + V[3 * i + 1] * l[1] + V[3 * i + 2] * l[2];
      if (v > bv) { bv = v; best = i; }
    }
    p[0] = V[3 * best]; p[1] = V[3 * best + 1]; p[2] = V[3 * best + 2];
  }
  out[0] = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + s.pos[0];
  out[1] = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + s.pos[1];
  out[2] = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + s.pos[2];
}

// support point of the Minkowski difference: v = a - b
template <class T>
struct SP { T v[3], a[3]; };

template <class T>
MM_HDN void support(const Shape<T>& s1, const Shape<T>& s2, const T* d, SP<T>& p) {
  T nd[3] = {-d[0], -d[1], -d[2]}, b[3];
  support1(s1, d, p.a);
  support1(s2, nd, b);
  for (int k = 0; k < 3; k++) p.v[k] = p.a[k] - b[k];
}

template <class T> MM_HD void sub3(T* r, const T* a, const T* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <class T> MM_HD void neg3(T* r, const T* a) { r[0] = -a[0]; r[1] = -a[1]; r[2] = -a[2]; }
template <class T> MM_HD void cpy3(T* r, const T* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }

// true if the shapes overlap; sx[0..3] then hold a tetrahedron around the origin
template <class T>
MM_HDN bool gjk(const Shape<T>& s1, const Shape<T>& s2, SP<T>* sx) {
  T dir[3] = {s2.pos[0] - s1.pos[0], s2.pos[1] - s1.pos[1], s2.pos[2] - s1.pos[2]};
  if (dot3(dir, dir) < (T)1e-20) { dir[0] = 1; dir[1] = 0; dir[2] = 0; }
  SP<T> a, b, c, d;
  support(s1, s2, dir, c);
  if (dot3(c.v, dir) < 0) return false;
  neg3(dir, c.v);
  if (dot3(dir, dir) < (T)1e-24) { dir[0] = 1; dir[1] = 0; dir[2] = 0; }
  support(s1, s2, dir, b);
  if (dot3(b.v, dir) < 0) return false;
  T bc[3], nb[3], t[3];
  sub3(bc, c.v, b.v);
  neg3(nb, b.v);
  cross3(t, bc, nb);
  cross3(dir, t, bc);
  if (dot3(dir, dir) < (T)1e-24) {
    const T ex[3] = {1, 0, 0}, ez[3] = {0, 0, 1};
    cross3(dir, bc, ex);
    if (dot3(dir, dir) < (T)1e-24) cross3(dir, bc, ez);
  }
  int n = 2;
  d = c;
  for (int it = 0; it < 64; it++) {
    support(s1, s2, dir, a);
    if (dot3(a.v, dir) < 0) return false;
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

// EPA scratch in the env's workspace
template <class T>
struct EpaMem {
  T* vert;   // [EPA_MAXV][6]  v(3), a(3)
  T* face;   // [EPA_MAXF][4]  n(3), d
  int* fidx; // [EPA_MAXF]     three vertex indices packed 10 bits each
  int* edge; // [EPA_MAXE]     two vertex indices packed 16 bits each
};

template <class T>
MM_HD void epa_mkface(const EpaMem<T>& m, int f, int ia, int ib, int ic) {
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
  if (d < 0) {  // wind outward
    F[0] = -n[0]; F[1] = -n[1]; F[2] = -n[2]; F[3] = -d;
    m.fidx[f] = ib | (ia << 10) | (ic << 20);
  } else {
    F[0] = n[0]; F[1] = n[1]; F[2] = n[2]; F[3] = d;
    m.fidx[f] = ia | (ib << 10) | (ic << 20);
  }
}

// Outputs contact position (mid witness), normal (shape1 -> shape2) and penetration depth.
template <class T>
MM_HDN bool epa(const Shape<T>& s1, const Shape<T>& s2, const SP<T>* sx, const EpaMem<T>& m, T* pos, T* nrm, T* depth) {
  int nv = 4, nf = 0;
  for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) { m.vert[6 * i + k] = sx[i].v[k]; m.vert[6 * i + 3 + k] = sx[i].a[k]; }
  // initial tetrahedron; the first four faces keep (p0, p2, p1) order when flipped, like the oracle
  const int tf[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
  for (int i = 0; i < 4; i++) {
    const T *a = m.vert + 6 * tf[i][0], *b = m.vert + 6 * tf[i][1], *c = m.vert + 6 * tf[i][2];
    T e1[3], e2[3], n[3];
    sub3(e1, b, a);
    sub3(e2, c, a);
    cross3(n, e1, e2);
    T l = tsqrt(dot3(n, n));
    T* F = m.face + 4 * nf;
    int i0 = tf[i][0], i1 = tf[i][1], i2 = tf[i][2];
    if ((double)l < 1e-30) { F[0] = F[1] = F[2] = 0; F[3] = (T)1e30; }
    else {
      T inv = (T)1 / l;
      n[0] *= inv; n[1] *= inv; n[2] *= inv;
      T d = dot3(n, a);
      if (d < 0) { F[0] = -n[0]; F[1] = -n[1]; F[2] = -n[2]; F[3] = -d; int t = i1; i1 = i2; i2 = t; }
      else { F[0] = n[0]; F[1] = n[1]; F[2] = n[2]; F[3] = d; }
    }
    m.fidx[nf] = i0 | (i1 << 10) | (i2 << 20);
    nf++;
  }
  for (int it = 0; it < EPA_MAXIT; it++) {
    int best = 0;
    for (int i = 1; i < nf; i++) if (m.face[4 * i + 3] < m.face[4 * best + 3]) best = i;
    T n[3] = {m.face[4 * best], m.face[4 * best + 1], m.face[4 * best + 2]};
    SP<T> p;
    support(s1, s2, n, p);
    T dist = dot3(p.v, n);
    if ((double)(dist - m.face[4 * best + 3]) < 1e-10 || nf >= EPA_MAXF - 8) break;
    int ip = nv++;
    for (int k = 0; k < 3; k++) { m.vert[6 * ip + k] = p.v[k]; m.vert[6 * ip + 3 + k] = p.a[k]; }
    int ne = 0;
    for (int i = 0; i < nf;) {
      int fi = m.fidx[i];
      int id[3] = {fi & 1023, (fi >> 10) & 1023, (fi >> 20) & 1023};
      const T* F = m.face + 4 * i;
      T r[3];
      sub3(r, p.v, m.vert + 6 * id[0]);
      if ((double)dot3(F, r) > 1e-14) {
        for (int e = 0; e < 3; e++) {
          int ea = id[e], eb = id[(e + 1) % 3];
          const T *va = m.vert + 6 * ea, *vb = m.vert + 6 * eb;
          bool found = false;
          for (int k = 0; k < ne; k++) {
            const T *ka = m.vert + 6 * (m.edge[k] & 0xFFFF), *kb = m.vert + 6 * (m.edge[k] >> 16);
            if (ka[0] == vb[0] && ka[1] == vb[1] && ka[2] == vb[2] && kb[0] == va[0] && kb[1] == va[1] && kb[2] == va[2]) {
              m.edge[k] = m.edge[--ne];
              found = true;
              break;
            }
          }
          if (!found && ne < EPA_MAXE) { m.edge[ne] = ea | (eb << 16); ne++; }
        }
        --nf;
        m.fidx[i] = m.fidx[nf];
        for (int k = 0; k < 4; k++) m.face[4 * i + k] = m.face[4 * nf + k];
      } else i++;
    }
    if (ne == 0) break;
    for (int k = 0; k < ne && nf < EPA_MAXF; k++) { epa_mkface(m, nf, m.edge[k] & 0xFFFF, m.edge[k] >> 16, ip); nf++; }
  }
  int best = 0;
  for (int i = 1; i < nf; i++) if (m.face[4 * i + 3] < m.face[4 * best + 3]) best = i;
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
