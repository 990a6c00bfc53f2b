/*
 * geometry_cvfem.h -- CVFEM dual-mesh geometry of one Tet4 / Wed6 / Pyr5
 * element: sub-control-volume volumes (one per node) and sub-control-surface
 * area vectors (one per integration point, each belonging to one element
 * edge), the per-element arithmetic of GeometryInteriorAlg
 * (src/ngp_algorithms/GeometryInteriorAlg.C:72-112, 165-225).
 *
 * Table driven: every topology is (i) a list of sub-points, each the average
 * of an ordered corner subset, (ii) sub-control volumes as hexahedra (the
 * pyramid's apex: a 10-point octohedron) of sub-points, (iii) sub-control
 * surfaces as quadrilaterals of sub-points (degenerate -- a repeated point --
 * on the pyramid's apex edges) and (iv) the (left, right) node pair of every
 * surface.  Corner order and association follow the reference so that results
 * agree to rounding:
 *   Tet4  src/master_element/Tet4CVFEM.C:243-343, 522-619; Tet4CVFEM.h:266
 *   Wed6  src/master_element/Wed6CVFEM.C:267-369, 541-647; Wed6CVFEM.h:268
 *   Pyr5  src/master_element/Pyr5CVFEM.C:348-572, 772-900; Pyr5CVFEM.h:300-301
 *   volumes / areas  include/master_element/Hex8GeometryFunctions.h:33-252
 *
 * NW_HD functions: compiled for the device by nw_geometry.cu and for the host
 * by the CPU walk-through of the test-suite (tests/emul).
 */
#ifndef NW_GEOMETRY_CVFEM_H
#define NW_GEOMETRY_CVFEM_H

#ifndef NW_HD
#if defined(__CUDACC__)
#define NW_HD __host__ __device__ __forceinline__
#else
#define NW_HD inline
#endif
#endif

namespace nw {
namespace geo {

enum Topology { TET4 = 0, WED6 = 1, PYR5 = 2 };

template <int T>
struct Traits;
template <>
struct Traits<TET4>
{
  enum { npe = 4, nSub = 15, nScv = 4, nScs = 6 };
};
template <>
struct Traits<WED6>
{
  enum { npe = 6, nSub = 21, nScv = 6, nScs = 9 };
};
template <>
struct Traits<PYR5>
{
  enum { npe = 5, nSub = 19, nScv = 5, nScs = 12 };
};

/* largest sizes over the three topologies */
enum { kMaxNpe = 6, kMaxSub = 21, kMaxScs = 12 };

/* one sub-point: r[0] corners averaged, r[1..] their order; r[0] == 0 is the
 * element centroid, accumulated corner by corner with the weight 1/npe */
NW_HD void
sub_point(
  const unsigned char* r, int npe, double wCentroid, const double c[][3],
  double* out)
{
  const double one3rd = 1.0 / 3.0;
  for (int d = 0; d < 3; ++d) {
    double a;
    switch (r[0]) {
    case 0:
      a = 0.0;
      for (int j = 0; j < npe; ++j)
        a = a + wCentroid * c[j][d];
      break;
    case 1:
      a = c[r[1]][d];
      break;
    case 2:
      a = 0.5 * (c[r[1]][d] + c[r[2]][d]);
      break;
    case 3:
      a = one3rd * (c[r[1]][d] + c[r[2]][d] + c[r[3]][d]);
      break;
    default:
      a = 0.25 * (c[r[1]][d] + c[r[2]][d] + c[r[3]][d] + c[r[4]][d]);
      break;
    }
    out[d] = a;
  }
}

template <int T>
NW_HD void sub_points(const double c[][3], double v[][3]);

template <>
NW_HD void
sub_points<TET4>(const double c[][3], double v[][3])
{
  const unsigned char R[15][5] = {
    {1, 0}, {1, 1}, {1, 2}, {1, 3},
    {2, 0, 1}, {2, 1, 2}, {2, 2, 0}, {3, 0, 1, 2},   /* face 0-1-2 */
    {2, 2, 3}, {2, 3, 1}, {3, 1, 2, 3},              /* face 1-2-3 */
    {2, 0, 3}, {3, 0, 2, 3},                         /* face 0-2-3 */
    {3, 0, 1, 3},                                    /* face 0-1-3 */
    {0}};
  for (int p = 0; p < 15; ++p)
    sub_point(R[p], 4, 0.25, c, v[p]);
}

template <>
NW_HD void
sub_points<WED6>(const double c[][3], double v[][3])
{
  const unsigned char R[21][5] = {
    {1, 0}, {1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5},
    {2, 0, 1}, {2, 1, 2}, {2, 2, 0}, {3, 0, 1, 2},   /* bottom triangle */
    {2, 3, 4}, {2, 4, 5}, {2, 5, 3}, {3, 3, 4, 5},   /* top triangle */
    {2, 1, 4}, {2, 0, 3}, {4, 0, 1, 4, 3},           /* quad 0-1-4-3 */
    {2, 2, 5}, {4, 1, 4, 5, 2},                      /* quad 1-4-5-2 */
    {4, 5, 3, 0, 2},                                 /* quad 5-3-0-2 */
    {0}};
  for (int p = 0; p < 21; ++p)
    sub_point(R[p], 6, 1.0 / 6.0, c, v[p]);
}

template <>
NW_HD void
sub_points<PYR5>(const double c[][3], double v[][3])
{
  const unsigned char R[19][5] = {
    {1, 0}, {1, 1}, {1, 2}, {1, 3}, {1, 4},
    {2, 0, 1}, {2, 1, 2}, {2, 2, 3}, {2, 3, 0}, {4, 0, 1, 2, 3}, /* base */
    {2, 1, 4}, {2, 4, 0}, {3, 0, 1, 4},
    {2, 2, 4}, {3, 1, 2, 4},
    {2, 3, 4}, {3, 3, 4, 2},
    {3, 0, 4, 3},
    {0}};
  for (int p = 0; p < 19; ++p)
    sub_point(R[p], 5, 0.2, c, v[p]);
}

/* one triangle's term of the divergence-theorem volume: (p + q + r) . ((q - p) x (r - p)) */
NW_HD void
tri_terms(const double* p, const double* q, const double* r, double* m, double* dxv)
{
  m[0] = p[0] + q[0] + r[0];
  m[1] = p[1] + q[1] + r[1];
  m[2] = p[2] + q[2] + r[2];
  dxv[0] = (q[1] - p[1]) * (r[2] - p[2]) - (r[1] - p[1]) * (q[2] - p[2]);
  dxv[1] = (r[0] - p[0]) * (q[2] - p[2]) - (q[0] - p[0]) * (r[2] - p[2]);
  dxv[2] = (q[0] - p[0]) * (r[1] - p[1]) - (r[0] - p[0]) * (q[1] - p[1]);
}

/* Grandy's 24-triangle hexahedron volume (hex_volume_grandy,
 * Hex8GeometryFunctions.h:83-158); bentTop: the variant whose top face is
 * split along its 5-7 diagonal (bhex_volume_grandy, :161-252) */
NW_HD double
grandy_volume(const double sc[8][3], bool bentTop)
{
  const unsigned char F[2][6][4] = {
    {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {1, 2, 6, 5}, {0, 4, 3, 7}},
    {{0, 1, 2, 3}, {5, 7, 5, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {1, 2, 6, 5}, {0, 3, 7, 4}}};
  const unsigned char TRI[24][3] = {
    {0, 8, 1},  {8, 2, 1},  {3, 2, 8},  {3, 8, 0},  {6, 9, 5},  {7, 9, 6},
    {4, 9, 7},  {4, 5, 9},  {10, 0, 1}, {5, 10, 1}, {4, 10, 5}, {4, 0, 10},
    {7, 6, 11}, {6, 2, 11}, {2, 3, 11}, {3, 7, 11}, {6, 12, 2}, {5, 12, 6},
    {5, 1, 12}, {1, 2, 12}, {0, 4, 13}, {4, 7, 13}, {7, 3, 13}, {3, 0, 13}};
  double cv[14][3];
  for (int n = 0; n < 8; ++n)
    for (int d = 0; d < 3; ++d)
      cv[n][d] = sc[n][d];
  const int b = bentTop ? 1 : 0;
  for (int k = 0; k < 6; ++k)
    for (int d = 0; d < 3; ++d) {
      if (bentTop && k == 1)
        cv[9][d] = 0.5 * (sc[5][d] + sc[7][d]);
      else
        cv[8 + k][d] = 0.25 * (sc[F[b][k][0]][d] + sc[F[b][k][1]][d] +
                               sc[F[b][k][2]][d] + sc[F[b][k][3]][d]);
    }
  double volume = 0.0;
  for (int k = 0; k < 24; ++k) {
    double m[3], dxv[3];
    tri_terms(cv[TRI[k][0]], cv[TRI[k][1]], cv[TRI[k][2]], m, dxv);
    volume += m[0] * dxv[0] + m[1] * dxv[1] + m[2] * dxv[2];
  }
  volume /= 18.0;
  return volume;
}

/* the pyramid's apex control volume: 10 vertices + the mid points of its four
 * non-planar faces, 24 triangles (octohedron_volume_by_triangle_facets and
 * polyhedral_volume_by_faces, Pyr5CVFEM.C:348-432) */
NW_HD double
octohedron_volume(const double vc[10][3])
{
  const unsigned char TRI[24][3] = {
    {1, 3, 10}, {2, 10, 3}, {2, 9, 10}, {10, 9, 1}, {4, 3, 11}, {3, 1, 11},
    {11, 1, 5}, {4, 11, 5}, {1, 12, 5}, {1, 7, 12}, {12, 7, 6}, {5, 12, 6},
    {9, 8, 13}, {13, 8, 7}, {13, 7, 1}, {9, 13, 1}, {4, 5, 0},  {5, 6, 0},
    {6, 7, 0},  {7, 8, 0},  {0, 8, 9},  {0, 9, 2},  {0, 2, 3},  {0, 3, 4}};
  double c[14][3];
  for (int j = 0; j < 10; ++j)
    for (int d = 0; d < 3; ++d)
      c[j][d] = vc[j][d];
  for (int d = 0; d < 3; ++d) {
    c[10][d] = 0.5 * (vc[3][d] + vc[9][d]);
    c[11][d] = 0.5 * (vc[3][d] + vc[5][d]);
    c[12][d] = 0.5 * (vc[5][d] + vc[7][d]);
    c[13][d] = 0.5 * (vc[7][d] + vc[9][d]);
  }
  double volume = 0.0;
  for (int k = 0; k < 24; ++k) {
    double m[3], dxv[3];
    tri_terms(c[TRI[k][0]], c[TRI[k][1]], c[TRI[k][2]], m, dxv);
    /* the reference writes the middle term as  - x1 * (-(dxv1))  and adds the
     * three terms to the running sum one at a time */
    volume = volume + m[0] * dxv[0] - m[1] * (-dxv[1]) + m[2] * dxv[2];
  }
  volume = volume / 18.0;
  return volume;
}

/* area vector of a quadrilateral facet: four-triangle fan about the mean of its
 * vertices (quad_area_by_triangulation, Hex8GeometryFunctions.h:33-81) */
NW_HD void
quad_area(const double* q0, const double* q1, const double* q2, const double* q3, double* a)
{
  const double* q[4] = {q0, q1, q2, q3};
  double xm[3], r1[3];
  for (int d = 0; d < 3; ++d) {
    xm[d] = 0.25 * (q0[d] + q1[d] + q2[d] + q3[d]);
    r1[d] = q0[d] - xm[d];
    a[d] = 0.0;
  }
  for (int it = 0; it < 4; ++it) {
    const double* qt = q[(it + 1) & 3];
    const double r2[3] = {qt[0] - xm[0], qt[1] - xm[1], qt[2] - xm[2]};
    a[0] += r1[1] * r2[2] - r2[1] * r1[2];
    a[1] += r1[2] * r2[0] - r2[2] * r1[0];
    a[2] += r1[0] * r2[1] - r2[0] * r1[1];
    r1[0] = r2[0];
    r1[1] = r2[1];
    r1[2] = r2[2];
  }
  a[0] *= 0.5;
  a[1] *= 0.5;
  a[2] *= 0.5;
}

/* volume of sub-control volume ip (its node is local node ip: ipNodeMap is the
 * identity for the three topologies) */
template <int T>
NW_HD double scv_volume(int ip, const double v[][3]);

NW_HD double
hex_of_sub_points(const unsigned char* t, const double v[][3], bool bentTop)
{
  double sc[8][3];
  for (int n = 0; n < 8; ++n)
    for (int d = 0; d < 3; ++d)
      sc[n][d] = v[t[n]][d];
  return grandy_volume(sc, bentTop);
}

template <>
NW_HD double
scv_volume<TET4>(int ip, const double v[][3])
{
  const unsigned char S[4][8] = {
    {0, 4, 7, 6, 11, 13, 14, 12}, {1, 5, 7, 4, 9, 10, 14, 13},
    {2, 6, 7, 5, 8, 12, 14, 10}, {3, 9, 13, 11, 8, 10, 14, 12}};
  return hex_of_sub_points(S[ip], v, false);
}

template <>
NW_HD double
scv_volume<WED6>(int ip, const double v[][3])
{
  const unsigned char S[6][8] = {
    {0, 15, 16, 6, 8, 19, 20, 9},    {9, 6, 1, 7, 20, 16, 14, 18},
    {8, 9, 7, 2, 19, 20, 18, 17},    {19, 15, 16, 20, 12, 3, 10, 13},
    {20, 16, 14, 18, 13, 10, 4, 11}, {19, 20, 18, 17, 12, 13, 11, 5}};
  return hex_of_sub_points(S[ip], v, false);
}

template <>
NW_HD double
scv_volume<PYR5>(int ip, const double v[][3])
{
  const unsigned char S[4][8] = {
    {0, 5, 9, 8, 11, 12, 18, 17}, {1, 6, 9, 5, 10, 14, 18, 12},
    {2, 7, 9, 6, 13, 16, 18, 14}, {3, 8, 9, 7, 15, 17, 18, 16}};
  if (ip < 4)
    return hex_of_sub_points(S[ip], v, true);
  const unsigned char A[10] = {4, 18, 15, 17, 11, 12, 10, 14, 13, 16};
  double oc[10][3];
  for (int n = 0; n < 10; ++n)
    for (int d = 0; d < 3; ++d)
      oc[n][d] = v[A[n]][d];
  return octohedron_volume(oc);
}

/* area vector of sub-control surface ip, pointing from its left to its right
 * node, and that node pair (lrscv) */
template <int T>
NW_HD void scs_area(int ip, const double v[][3], double* a);
template <int T>
NW_HD void scs_nodes(int ip, int* l, int* r);

template <>
NW_HD void
scs_area<TET4>(int ip, const double v[][3], double* a)
{
  const unsigned char Q[6][4] = {{4, 7, 14, 13},  {7, 14, 10, 5}, {6, 12, 14, 7},
                                 {11, 13, 14, 12}, {13, 9, 10, 14}, {10, 8, 12, 14}};
  quad_area(v[Q[ip][0]], v[Q[ip][1]], v[Q[ip][2]], v[Q[ip][3]], a);
}
template <>
NW_HD void
scs_nodes<TET4>(int ip, int* l, int* r)
{
  const unsigned char LR[12] = {0, 1, 1, 2, 0, 2, 0, 3, 1, 3, 2, 3};
  *l = LR[2 * ip];
  *r = LR[2 * ip + 1];
}

template <>
NW_HD void
scs_area<WED6>(int ip, const double v[][3], double* a)
{
  const unsigned char Q[9][4] = {
    {6, 9, 20, 16},   {7, 9, 20, 18},   {9, 8, 19, 20},  {10, 16, 20, 13}, {13, 11, 18, 20},
    {12, 13, 20, 19}, {15, 16, 20, 19}, {16, 14, 18, 20}, {19, 20, 18, 17}};
  quad_area(v[Q[ip][0]], v[Q[ip][1]], v[Q[ip][2]], v[Q[ip][3]], a);
}
template <>
NW_HD void
scs_nodes<WED6>(int ip, int* l, int* r)
{
  const unsigned char LR[18] = {0, 1, 1, 2, 0, 2, 3, 4, 4, 5, 3, 5, 0, 3, 1, 4, 2, 5};
  *l = LR[2 * ip];
  *r = LR[2 * ip + 1];
}

template <>
NW_HD void
scs_area<PYR5>(int ip, const double v[][3], double* a)
{
  /* surfaces 4..11: two facets (inner, outer) per apex edge */
  const unsigned char Q[12][4] = {
    {5, 9, 18, 12},   {6, 9, 18, 14},   {7, 9, 18, 16},   {8, 17, 18, 9},
    {12, 12, 18, 17}, {11, 12, 12, 17}, {14, 14, 18, 12}, {10, 14, 14, 12},
    {16, 16, 18, 14}, {13, 16, 16, 14}, {17, 17, 18, 16}, {15, 17, 17, 16}};
  quad_area(v[Q[ip][0]], v[Q[ip][1]], v[Q[ip][2]], v[Q[ip][3]], a);
}
template <>
NW_HD void
scs_nodes<PYR5>(int ip, int* l, int* r)
{
  const unsigned char LR[24] = {0, 1, 1, 2, 2, 3, 0, 3, 0, 4, 0, 4,
                                1, 4, 1, 4, 2, 4, 2, 4, 3, 4, 3, 4};
  *l = LR[2 * ip];
  *r = LR[2 * ip + 1];
}

} // namespace geo
} // namespace nw

#endif
