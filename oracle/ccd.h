// ORACLE - TEST INFRASTRUCTURE ONLY (see engine.h).
// General convex-convex penetration query (GJK intersection test + EPA expansion) used for every
// geom pair that involves a collision mesh (convex hull) or a cylinder.  Restates what MuJoCo's
// native convex collision pipeline returns for such pairs with `multiccd` off:
// one contact, normal from geom1 to geom2, position midway between the two witness points,
// dist = -penetration depth.
#pragma once
#include <cmath>
#include <cstring>

namespace ccd {

// high-water marks of the EPA polytope (sizing aid for the GPU workspace; not part of the algorithm)
struct Marks { int max_faces, max_iters; long long calls, iters, gjk_calls, gjk_iters, gjk_hits, gjk_maxed; };
inline Marks& marks() { static Marks m = {0, 0, 0, 0, 0, 0, 0, 0}; return m; }

struct Shape {
  int type;            // 5 cylinder, 6 box, 7 convex mesh
  const double* pos;   // world position (3)
  const double* mat;   // world orientation (row-major 3x3)
  const double* size;  // box half sizes / cylinder (radius, half height)
  const double* verts; // mesh hull vertices, geom frame (3*nvert)
  int nvert;
};

struct V3 { double x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

static inline V3 support1(const Shape& s, V3 d) {
  const double* R = s.mat;
  // direction in the geom frame
  double lx = R[0] * d.x + R[3] * d.y + R[6] * d.z;
  double ly = R[1] * d.x + R[4] * d.y + R[7] * d.z;
  double lz = R[2] * d.x + R[5] * d.y + R[8] * d.z;
  double px, py, pz;
  if (s.type == 6) {
    px = lx >= 0 ? s.size[0] : -s.size[0];
    py = ly >= 0 ? s.size[1] : -s.size[1];
    pz = lz >= 0 ? s.size[2] : -s.size[2];
  } else if (s.type == 5) {
    double n = std::sqrt(lx * lx + ly * ly);
    if (n > 1e-14) { px = lx / n * s.size[0]; py = ly / n * s.size[0]; } else { px = py = 0; }
    pz = lz >= 0 ? s.size[1] : -s.size[1];
  } else {
    int best = 0;
    double bv = -1e300;
    for (int i = 0; i < s.nvert; i++) {
      double v = s.verts[3 * i] * lx + s.verts[3 * i + 1] * ly + s.verts[3 * i + 2] * lz;
      if (v > bv) { bv = v; best = i; }
    }
    px = s.verts[3 * best]; py = s.verts[3 * best + 1]; pz = s.verts[3 * best + 2];
  }
  return {R[0] * px + R[1] * py + R[2] * pz + s.pos[0], R[3] * px + R[4] * py + R[5] * pz + s.pos[1],
          R[6] * px + R[7] * py + R[8] * pz + s.pos[2]};
}

struct SP { V3 v, a, b; };  // v = a - b
static inline SP support(const Shape& s1, const Shape& s2, V3 d) {
  SP p;
  p.a = support1(s1, d);
  p.b = support1(s2, neg(d));
  p.v = p.a - p.b;
  return p;
}

// GJK: true if the shapes overlap; on success simplex[0..3] is a tetrahedron enclosing the origin
static inline bool gjk(const Shape& s1, const Shape& s2, SP* sx) {
  V3 dir = {s2.pos[0] - s1.pos[0], s2.pos[1] - s1.pos[1], s2.pos[2] - s1.pos[2]};
  if (dot(dir, dir) < 1e-20) dir = {1, 0, 0};
  SP a, b, c, d;
  c = support(s1, s2, dir);
  if (dot(c.v, dir) < 0) return false;
  dir = neg(c.v);
  if (dot(dir, dir) < 1e-24) dir = {1, 0, 0};
  b = support(s1, s2, dir);
  if (dot(b.v, dir) < 0) return false;
  V3 bc = c.v - b.v;
  dir = cross(cross(bc, neg(b.v)), bc);
  if (dot(dir, dir) < 1e-24) {  // origin on the segment: any perpendicular
    dir = cross(bc, V3{1, 0, 0});
    if (dot(dir, dir) < 1e-24) dir = cross(bc, V3{0, 0, 1});
  }
  int n = 2;
  marks().gjk_calls++;
  for (int it = 0; it < 64; it++) {
    marks().gjk_iters++;
    if (it == 63) marks().gjk_maxed++;
    a = support(s1, s2, dir);
    if (dot(a.v, dir) < 0) return false;
    if (n == 2) {
      // triangle a,b,c
      V3 ao = neg(a.v), ab = b.v - a.v, ac = c.v - a.v, nrm = cross(ab, ac);
      if (dot(cross(ab, nrm), ao) > 0) {  // outside edge ab
        c = a; dir = cross(cross(ab, ao), ab);
        if (dot(dir, dir) < 1e-24) dir = nrm;
        continue;
      }
      if (dot(cross(nrm, ac), ao) > 0) {  // outside edge ac
        b = a; dir = cross(cross(ac, ao), ac);
        if (dot(dir, dir) < 1e-24) dir = nrm;
        continue;
      }
      n = 3;
      if (dot(nrm, ao) > 0) { d = c; c = b; b = a; dir = nrm; }
      else { d = b; b = a; dir = neg(nrm); }
      continue;
    }
    // tetrahedron a (new), b, c, d with b,c,d wound so that the origin is on the a side
    V3 ao = neg(a.v), ab = b.v - a.v, ac = c.v - a.v, ad = d.v - a.v;
    V3 abc = cross(ab, ac), acd = cross(ac, ad), adb = cross(ad, ab);
    if (dot(abc, ao) > 0) { d = c; c = b; b = a; dir = abc; continue; }
    if (dot(acd, ao) > 0) { b = a; dir = acd; continue; }
    if (dot(adb, ao) > 0) { c = d; d = b; b = a; dir = adb; continue; }
    sx[0] = a; sx[1] = b; sx[2] = c; sx[3] = d;
    marks().gjk_hits++;
    return true;
  }
  return false;
}

// EPA on the Minkowski difference.  Outputs contact position (mid witness), normal (shape1 ->
// shape2) and penetration depth.
static inline bool epa(const Shape& s1, const Shape& s2, const SP* sx, double* pos, double* nrm, double* depth) {
  constexpr int MAXF = 256, MAXE = 128, MAXIT = 128;
  struct Face { SP p[3]; V3 n; double d; };
  static thread_local Face faces[MAXF];
  int nf = 0;
  auto mk = [&](const SP& a, const SP& b, const SP& c) {
    Face f;
    f.p[0] = a; f.p[1] = b; f.p[2] = c;
    V3 n = cross(b.v - a.v, c.v - a.v);
    double l = std::sqrt(dot(n, n));
    if (l < 1e-30) { f.n = {0, 0, 0}; f.d = 1e300; }
    else { f.n = n * (1.0 / l); f.d = dot(f.n, a.v); }
    return f;
  };
  // tetra from gjk: a,b,c,d  (faces wound outward)
  faces[nf++] = mk(sx[0], sx[1], sx[2]);
  faces[nf++] = mk(sx[0], sx[2], sx[3]);
  faces[nf++] = mk(sx[0], sx[3], sx[1]);
  faces[nf++] = mk(sx[1], sx[3], sx[2]);
  for (int i = 0; i < 4; i++)
    if (faces[i].d < 0) {  // fix winding if needed
      SP t = faces[i].p[1]; faces[i].p[1] = faces[i].p[2]; faces[i].p[2] = t;
      faces[i].n = neg(faces[i].n); faces[i].d = -faces[i].d;
    }
  int best = 0;
  for (int it = 0; it < MAXIT; it++) {
    best = 0;
    for (int i = 1; i < nf; i++) if (faces[i].d < faces[best].d) best = i;
    V3 n = faces[best].n;
    SP p = support(s1, s2, n);
    double dist = dot(p.v, n);
    if (nf > marks().max_faces) marks().max_faces = nf;
    if (it > marks().max_iters) marks().max_iters = it;
    marks().iters++;
    if (dist - faces[best].d < 1e-10 || nf >= MAXF - 8) break;
    // remove faces visible from p, collect the horizon
    struct Edge { SP a, b; };
    Edge edges[MAXE];
    int ne = 0;
    for (int i = 0; i < nf;) {
      if (dot(faces[i].n, p.v - faces[i].p[0].v) > 1e-14) {
        for (int e = 0; e < 3; e++) {
          SP ea = faces[i].p[e], eb = faces[i].p[(e + 1) % 3];
          bool found = false;
          for (int k = 0; k < ne; k++) {
            // shared edges appear reversed in the neighbouring face
            if (edges[k].a.v.x == eb.v.x && edges[k].a.v.y == eb.v.y && edges[k].a.v.z == eb.v.z &&
                edges[k].b.v.x == ea.v.x && edges[k].b.v.y == ea.v.y && edges[k].b.v.z == ea.v.z) {
              edges[k] = edges[--ne];
              found = true;
              break;
            }
          }
          if (!found && ne < MAXE) { edges[ne].a = ea; edges[ne].b = eb; ne++; }
        }
        faces[i] = faces[--nf];
      } else i++;
    }
    if (ne == 0) break;
    for (int k = 0; k < ne; k++) {
      Face f = mk(edges[k].a, edges[k].b, p);
      if (f.d < 0) { SP t = f.p[0]; f.p[0] = f.p[1]; f.p[1] = t; f.n = neg(f.n); f.d = -f.d; }
      faces[nf++] = f;
    }
  }
  marks().calls++;
  best = 0;
  for (int i = 1; i < nf; i++) if (faces[i].d < faces[best].d) best = i;
  const Face& f = faces[best];
  if (!(f.d < 1e299)) return false;
  // barycentric coordinates of the origin's projection onto the closest face
  V3 pr = f.n * f.d;
  V3 v0 = f.p[1].v - f.p[0].v, v1 = f.p[2].v - f.p[0].v, v2 = pr - f.p[0].v;
  double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
  double den = d00 * d11 - d01 * d01;
  double v = 0, w = 0;
  if (std::fabs(den) > 1e-30) { v = (d11 * d20 - d01 * d21) / den; w = (d00 * d21 - d01 * d20) / den; }
  double u = 1 - v - w;
  V3 wa = f.p[0].a * u + f.p[1].a * v + f.p[2].a * w;
  V3 wb = f.p[0].b * u + f.p[1].b * v + f.p[2].b * w;
  pos[0] = 0.5 * (wa.x + wb.x); pos[1] = 0.5 * (wa.y + wb.y); pos[2] = 0.5 * (wa.z + wb.z);
  nrm[0] = f.n.x; nrm[1] = f.n.y; nrm[2] = f.n.z;
  *depth = f.d;
  return f.d > 0;
}

static inline bool penetration(const Shape& s1, const Shape& s2, double* pos, double* nrm, double* depth) {
  SP sx[4];
  if (!gjk(s1, s2, sx)) return false;
  return epa(s1, s2, sx, pos, nrm, depth);
}

}  // namespace ccd
