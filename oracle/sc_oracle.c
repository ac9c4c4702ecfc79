/* TEST INFRASTRUCTURE -- CPU oracle for the scOOP pair-energy hot path. NOT part of the product.
 * See sc_oracle.h for scope and parity status. Every function cites the reference lines it restates
 * (paths relative to /root/reference/scOOP/). Arithmetic is written in the reference's operation
 * order; compile with -ffp-contract=off and without -ffast-math so that it is IEEE-reproducible.
 */
#include "sc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PIH 1.57079632679489661923132169163975 /* structures/macros.h:66 */

typedef struct { double x, y, z; } vec3;

static inline vec3 V(double x, double y, double z) { vec3 v; v.x = x; v.y = y; v.z = z; return v; }
static inline vec3 ld(const double* p) { return V(p[0], p[1], p[2]); }
/* DOT macro, structures/macros.h:110 : (ax*bx + ay*by) + az*bz */
static inline double dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* Vector::size, structures/Vector.h:48-50 (pow(x,2) is exactly x*x) */
static inline double vsize(vec3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
/* operator*(double, Vector), structures/Vector.h:317-319 */
static inline vec3 scal(double s, vec3 v) { return V(v.x * s, v.y * s, v.z * s); }
/* vecCrossProduct, mc/math_calc.h:40-42 */
static inline vec3 cross(vec3 A, vec3 B) {
    return V(A.y * B.z - A.z * B.y, -A.x * B.z + A.z * B.x, A.x * B.y - A.y * B.x);
}
/* Vector::perpProject, structures/Vector.h:233-244 */
static inline vec3 perp_project(vec3 a, vec3 B) {
    double dp = dot(a, B);
    return V(a.x - B.x * dp, a.y - B.y * dp, a.z - B.z * dp);
}

/* anInt, mc/math_calc.h:15-26: add 1.5*2^52 and read the low 32 bits as a signed int */
static inline double an_int(double arg) {
    arg += 6755399441055744.0;
    int lo;
    memcpy(&lo, &arg, sizeof(int)); /* little endian: low word first */
    return (double)lo;
}

/* Cuboid::image, structures/geometry.h:110-128 */
static inline vec3 image(const double box[3], vec3 r1, vec3 r2) {
    vec3 r = V(r1.x - r2.x, r1.y - r2.y, r1.z - r2.z);
    r.x = box[0] * (r.x - an_int(r.x));
    r.y = box[1] * (r.y - an_int(r.y));
    r.z = box[2] * (r.z - an_int(r.z));
    return r;
}

void sco_image(const double box[3], const double r1[3], const double r2[3], double out[3]) {
    vec3 r = image(box, ld(r1), ld(r2));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* EPatch::minDistSegments, mc/paire.cpp:85-241 */
static vec3 min_dist_segments(vec3 segA, vec3 segB, double halfl1, double halfl2, vec3 r_cm) {
    vec3 u, v, w, vec;
    double a, b, c, d, e, D, sc, sN, sD, tc, tN, tD;
    int paralel = 0;

    u = scal(2.0 * halfl1, segA);
    v = scal(2.0 * halfl2, segB);
    w.x = segB.x * halfl2 - segA.x * halfl1 - r_cm.x;
    w.y = segB.y * halfl2 - segA.y * halfl1 - r_cm.y;
    w.z = segB.z * halfl2 - segA.z * halfl1 - r_cm.z;

    a = dot(u, u);
    b = dot(u, v);
    c = dot(v, v);
    d = dot(u, w);
    e = dot(v, w);
    D = a * c - b * b;
    sc = D; sN = D; sD = D;
    tc = D; tN = D; tD = D;

    if (D < 0.00000001) {
        paralel = 1;
        sN = 0.0;
        sD = 1.0;
        tN = e;
        tD = c;
    } else {
        sN = (b * e - c * d);
        tN = (a * e - b * d);
        if (sN < 0.0) {
            sN = 0.0;
            tN = e;
            tD = c;
        } else if (sN > sD) {
            sN = sD;
            tN = e + b;
            tD = c;
        }
    }

    if (tN < 0.0) {
        tN = 0.0;
        if (-d < 0.0) sN = 0.0;
        else if (-d > a) sN = sD;
        else { sN = -d; sD = a; }
    } else if (tN > tD) {
        tN = tD;
        if ((-d + b) < 0.0) sN = 0;
        else if ((-d + b) > a) sN = sD;
        else { sN = (-d + b); sD = a; }
    }
    if (fabs(sN) < 0.00000001) sc = 0.0; else sc = sN / sD;
    if (fabs(tN) < 0.00000001) tc = 0.0; else tc = tN / tD;

    vec.x = u.x * sc + w.x - v.x * tc;
    vec.y = u.y * sc + w.y - v.y * tc;
    vec.z = u.z * sc + w.z - v.z * tc;

    if (paralel) { /* second pass with the roles of the segments swapped, paire.cpp:170-238 */
        vec3 vec2;
        w.x = segA.x * halfl1 - segB.x * halfl2 + r_cm.x;
        w.y = segA.y * halfl1 - segB.y * halfl2 + r_cm.y;
        w.z = segA.z * halfl1 - segB.z * halfl2 + r_cm.z;
        d = dot(v, w);
        e = dot(u, w);
        D = a * c - b * b;
        sc = D; sN = D; sD = D;
        tc = D; tN = D; tD = D;
        if (D < 0.00000001) {
            sN = 0.0;
            sD = 1.0;
            tN = e;
            tD = a;
        }
        if (tN < 0.0) {
            tN = 0.0;
            if (-d < 0.0) sN = 0.0;
            else if (-d > c) sN = sD;
            else { sN = -d; sD = c; }
        } else if (tN > tD) {
            tN = tD;
            if ((-d + b) < 0.0) sN = 0;
            else if ((-d + b) > c) sN = sD;
            else { sN = (-d + b); sD = c; }
        }
        if (fabs(sN) < 0.00000001) sc = 0.0; else sc = sN / sD;
        if (fabs(tN) < 0.00000001) tc = 0.0; else tc = tN / tD;
        vec2.x = v.x * sc + w.x - u.x * tc;
        vec2.y = v.y * sc + w.y - u.y * tc;
        vec2.z = v.z * sc + w.z - u.z * tc;
        if (dot(vec2, vec2) < dot(vec, vec)) return vec2;
    }
    return vec;
}

void sco_min_dist_segments(const double segA[3], const double segB[3], double halfl1, double halfl2,
                           const double r_cm[3], double out[3]) {
    vec3 r = min_dist_segments(ld(segA), ld(segB), halfl1, halfl2, ld(r_cm));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* fanglScale, mc/paire.h:12-17 */
static inline double fangl_scale(double a, double pcangl, double pcanglsw) {
    if (a <= pcanglsw) return 0.0;
    return (a >= pcangl) ? 1.0 : (0.5 - ((pcanglsw + pcangl) * 0.5 - a) / (pcangl - pcanglsw));
}

typedef struct { vec3 dir; vec3 sides[2]; } patch_t; /* Patch, structures/particle.h:10-20 */

/* EPatch::testIntrPatch, mc/paire.h:86-101 */
static inline void test_intr_patch(vec3 dir, vec3 patchdir, vec3 vec, double cospatch, double ti, double in[2]) {
    vec = perp_project(vec, dir);
    if (dot(patchdir, vec) >= cospatch * vsize(vec)) {
        if (in[0] == 0) { in[0] = ti; return; }
        if (in[1] == 0 && in[0] != ti) { in[1] = ti; return; }
    }
}

/* EPatch::scToInfiIntr, mc/paire.h:103-119 */
static inline void sc_to_infi_intr(vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, vec3 r_cm, double pcanglsw,
                                   double halfl1, double halfl2, double in[2], double x1) {
    if ((x1 >= halfl2) || (x1 <= -halfl2)) {
        ;
    } else {
        vec3 vec1 = V(p2Dir.x * x1 - r_cm.x, p2Dir.y * x1 - r_cm.y, p2Dir.z * x1 - r_cm.z);
        double e = dot(p1Dir, vec1);
        if ((e >= halfl1) || (e <= -halfl1)) ;
        else test_intr_patch(p1Dir, p1P->dir, vec1, pcanglsw, x1, in);
    }
}

/* EPatch::testIntrAtC, mc/paire.h:121-148 */
static void test_intr_at_c(vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, vec3 r_cm, double pcanglsw, double rcutSq,
                           double halfl1, double halfl2, double in[2]) {
    vec3 vec1 = cross(scal(-1.0, r_cm), p1Dir);
    vec3 vec2 = cross(p2Dir, p1Dir);
    double a = dot(vec2, vec2);
    double b = 2 * dot(vec1, vec2);
    double c = -rcutSq + dot(vec1, vec1);
    double d = b * b - 4 * a * c;
    if (d >= 0) {
        double x1;
        d = sqrt(d);
        a = 0.5 / a;
        x1 = (-b + d) * a;
        sc_to_infi_intr(p1Dir, p2Dir, p1P, r_cm, pcanglsw, halfl1, halfl2, in, x1);
        if (d > 0) {
            x1 = (-b - d) * a;
            sc_to_infi_intr(p1Dir, p2Dir, p1P, r_cm, pcanglsw, halfl1, halfl2, in, x1);
        }
    }
}

/* EPatch::findIntersectPlaneUni, mc/paire.h:150-181 */
static int find_intersect_plane_uni(vec3 dirA, vec3 dirB, double halfl, vec3 r_cm, vec3 w_vec, double cospatch,
                                    double* ti, double* c, double* d) {
    vec3 nplane = cross(dirA, w_vec);
    double a = dot(nplane, dirB);
    *c = 1.0; *d = 1.0;
    if (a == 0.0) return 0;
    *ti = dot(nplane, r_cm) / a;
    if ((*ti > halfl) || (*ti < -halfl)) return 0;
    {
        vec3 d_vec = V(*ti * dirB.x - r_cm.x, *ti * dirB.y - r_cm.y, *ti * dirB.z - r_cm.z);
        *c = dot(d_vec, w_vec);
        if (*c * cospatch < 0) return 0;
        *d = fabs(dot(d_vec, dirA)) - halfl;
        return 1;
    }
}

/* Psc::scToEndSpIntr, mc/paire.h:588-602 */
static inline void sc_to_end_sp_intr(vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, vec3 r_cm, double pcanglsw,
                                     double halfl1, double halfl2, double in[2], double x1) {
    if ((x1 >= halfl2) || (x1 <= -halfl2)) {
        ;
    } else {
        vec3 vec1 = V(p2Dir.x * x1 - r_cm.x, p2Dir.y * x1 - r_cm.y, p2Dir.z * x1 - r_cm.z);
        double e = dot(p1Dir, vec1);
        if ((e >= halfl1) || (e <= -halfl1)) test_intr_patch(p1Dir, p1P->dir, vec1, pcanglsw, x1, in);
    }
}

/* Psc::calcIntersections, mc/paire.h:605-623 */
static inline void calc_intersections(vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, vec3 r_cm, double in[2],
                                      double pcanglsw, double halfl1, double halfl2, double b, double c) {
    double d = b * b - 4 * c;
    if (d >= 0) {
        d = sqrt(d);
        c = (-b + d) * 0.5;
        sc_to_end_sp_intr(p1Dir, p2Dir, p1P, r_cm, pcanglsw, halfl1, halfl2, in, c);
        if (d > 0) {
            c = (-b - d) * 0.5;
            sc_to_end_sp_intr(p1Dir, p2Dir, p1P, r_cm, pcanglsw, halfl1, halfl2, in, c);
        }
    }
}

/* Psc::testIntrA, mc/paire.h:625-652 */
static inline void test_intr_a(vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, vec3 r_cm, double pcanglsw, double rcutSq,
                               double halfl1, double halfl2, double in[2]) {
    vec3 vec1 = V(p2Dir.x * halfl2 - r_cm.x, p2Dir.y * halfl2 - r_cm.y, p2Dir.z * halfl2 - r_cm.z);
    double a = dot(vec1, p1Dir);
    vec3 vec2 = V(vec1.x - p1Dir.x * a, vec1.y - p1Dir.y * a, vec1.z - p1Dir.z * a);
    double b = dot(vec2, vec2);
    double d = fabs(a) - halfl1;
    double c;
    if (d <= 0) c = b; else c = d * d + b;
    if (c < rcutSq) test_intr_patch(p1Dir, p1P->dir, vec1, pcanglsw, halfl2, in);
}

/* Psc::pscIntersect, mc/paire.h:499-583 */
static int psc_intersect(vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, vec3 r_cm, double* in1, double* in2,
                         double pcanglsw, double rcutSq, double halfl1, double halfl2) {
    double c, d, ti, disti;
    double in[2] = {0, 0};

    if (find_intersect_plane_uni(p1Dir, p2Dir, halfl2, r_cm, p1P->sides[0], pcanglsw, &ti, &c, &d)) {
        if (d <= 0) disti = c * c; else disti = d * d + c * c;
        if (disti <= rcutSq) in[0] = ti;
    }
    if (find_intersect_plane_uni(p1Dir, p2Dir, halfl2, r_cm, p1P->sides[1], pcanglsw, &ti, &c, &d)) {
        if (d <= 0) disti = c * c; else disti = d * d + c * c;
        if (disti <= rcutSq) {
            if (in[0] == 0.0) in[0] = ti;
            else if (ti != in[0]) in[1] = ti;
        }
    }
    if (in[1] != 0.0) { *in1 = in[0]; *in2 = in[1]; return 2; }
    test_intr_at_c(p1Dir, p2Dir, p1P, r_cm, pcanglsw, rcutSq, halfl1, halfl2, in);

    if (in[1] == 0.0) {
        vec3 vec1 = V(p1Dir.x * halfl1 - r_cm.x, p1Dir.y * halfl1 - r_cm.y, p1Dir.z * halfl1 - r_cm.z);
        vec3 vec2 = V(-p1Dir.x * halfl1 - r_cm.x, -p1Dir.y * halfl1 - r_cm.y, -p1Dir.z * halfl1 - r_cm.z);
        calc_intersections(p1Dir, p2Dir, p1P, r_cm, in, pcanglsw, halfl1, halfl2, 2.0 * dot(vec1, p2Dir), dot(vec1, vec1) - rcutSq);
        calc_intersections(p1Dir, p2Dir, p1P, r_cm, in, pcanglsw, halfl1, halfl2, 2.0 * dot(vec2, p2Dir), dot(vec2, vec2) - rcutSq);
    } else { *in1 = in[0]; *in2 = in[1]; return 2; }

    if (in[1] == 0.0) {
        test_intr_a(p1Dir, p2Dir, p1P, r_cm, pcanglsw, rcutSq, halfl1, halfl2, in);
        if (in[1] == 0.0) test_intr_a(p1Dir, p2Dir, p1P, r_cm, pcanglsw, rcutSq, halfl1, -halfl2, in);
    }
    *in1 = in[0]; *in2 = in[1];
    return (in[1] == 0.0) ? 0 : 2;
}

/* CPsc::cpscIntersect, mc/paire.h:687-826 */
static int cpsc_intersect(vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, vec3 r_cm, double* in1, double* in2,
                          double pcanglsw, double rcutSq, double halfl1, double halfl2) {
    double a, b, c, d, x1, x2, disti = 0.0, ti;
    vec3 vec1, vec2;
    double in[2] = {0, 0};

    if (find_intersect_plane_uni(p1Dir, p2Dir, halfl2, r_cm, p1P->sides[0], pcanglsw, &ti, &c, &d)) {
        if (d <= 0) {
            disti = c * c;
            if (disti <= rcutSq) in[0] = ti;
        }
    }
    if (find_intersect_plane_uni(p1Dir, p2Dir, halfl2, r_cm, p1P->sides[1], pcanglsw, &ti, &c, &d)) {
        if (d <= 0) {
            disti = c * c;
            if (disti <= rcutSq) {
                if (in[0] == 0.0) in[0] = ti;
                else if (ti != in[0]) in[1] = ti;
            }
        }
    }
    if (in[1] != 0.0) { *in1 = in[0]; *in2 = in[1]; return 2; }
    test_intr_at_c(p1Dir, p2Dir, p1P, r_cm, pcanglsw, rcutSq, halfl1, halfl2, in);

    if (in[1] == 0.0) { /* end plates, paire.h:736-777 */
        a = dot(p1Dir, p2Dir);
        if (a == 0.0) {
            ;
        } else {
            vec1 = V(r_cm.x + halfl1 * p1Dir.x, r_cm.y + halfl1 * p1Dir.y, r_cm.z + halfl1 * p1Dir.z);
            x1 = dot(p1Dir, vec1) / a;
            if ((x1 > halfl2) || (x1 < -halfl2)) ;
            else {
                vec2 = V(x1 * p2Dir.x - vec1.x, x1 * p2Dir.y - vec1.y, x1 * p2Dir.z - vec1.z);
                b = dot(vec2, vec2);
                if (b > rcutSq) ;
                else test_intr_patch(p1Dir, p1P->dir, vec2, pcanglsw, x1, in);
            }
            vec1 = V(r_cm.x - halfl1 * p1Dir.x, r_cm.y - halfl1 * p1Dir.y, r_cm.z - halfl1 * p1Dir.z);
            x2 = dot(p1Dir, vec1) / a;
            if ((x2 > halfl2) || (x2 < -halfl2)) ;
            else {
                vec2 = V(x2 * p2Dir.x - vec1.x, x2 * p2Dir.y - vec1.y, x2 * p2Dir.z - vec1.z);
                b = dot(vec2, vec2);
                if (b > rcutSq) ;
                else test_intr_patch(p1Dir, p1P->dir, vec2, pcanglsw, x2, in);
            }
        }
    } else { *in1 = in[0]; *in2 = in[1]; return 2; }

    if (in[1] == 0.0) { /* rod-2 end points inside the cylindrical part, paire.h:786-821 */
        vec1 = V(p2Dir.x * halfl2 - r_cm.x, p2Dir.y * halfl2 - r_cm.y, p2Dir.z * halfl2 - r_cm.z);
        a = dot(vec1, p1Dir);
        vec2 = V(vec1.x - p1Dir.x * a, vec1.y - p1Dir.y * a, vec1.z - p1Dir.z * a);
        b = dot(vec2, vec2);
        d = fabs(a) - halfl1;
        if (d <= 0) {
            if (b < rcutSq) test_intr_patch(p1Dir, p1P->dir, vec1, pcanglsw, halfl2, in);
        }
        if (in[1] == 0.0) {
            vec1 = V(-p2Dir.x * halfl2 - r_cm.x, -p2Dir.y * halfl2 - r_cm.y, -p2Dir.z * halfl2 - r_cm.z);
            a = dot(vec1, p1Dir);
            vec2 = V(vec1.x - p1Dir.x * a, vec1.y - p1Dir.y * a, vec1.z - p1Dir.z * a);
            b = dot(vec2, vec2);
            d = fabs(a) - halfl1;
            if (d <= 0) {
                if (b < rcutSq) test_intr_patch(p1Dir, p1P->dir, vec1, pcanglsw, -1.0 * halfl2, in);
            }
        }
    }
    *in1 = in[0]; *in2 = in[1];
    return (in[1] == 0.0) ? 0 : 2;
}

/* EPatch::scparallel, mc/paire.h:77-84 */
static inline double scparallel(double epsilonparallel, vec3 dir1, vec3 dir2) {
    double cosa = dot(dir1, dir2);
    if ((epsilonparallel > 0 && cosa > 0) || (epsilonparallel < 0 && cosa < 0)) return 1.0 + epsilonparallel * cosa;
    return 1.0;
}

/* EPatch::atrE, mc/paire.cpp:243-301 */
static double atr_e(const sco_iaparam* ia, vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, const patch_t* p2P, vec3 r_cm,
                    int patchnum1, int patchnum2, double S1, double S2, double T1, double T2) {
    vec3 vec1, vec2, vec_intrs, vec_mindist;
    double atrenergy, ndist, a, paral, f1, f2;
    double v1 = fabs(S1 - S2);
    double v2 = fabs(T1 - T2);
    double f0 = 0.5 * (v1 + v2);

    vec1 = scal((S1 + S2) * 0.5, p1Dir);
    vec2 = scal((T1 + T2) * 0.5, p2Dir);
    vec_intrs.x = vec2.x - vec1.x - r_cm.x;
    vec_intrs.y = vec2.y - vec1.y - r_cm.y;
    vec_intrs.z = vec2.z - vec1.z - r_cm.z;

    vec_mindist = min_dist_segments(p1Dir, p2Dir, v1, v2, vec_intrs);
    ndist = sqrt(dot(vec_mindist, vec_mindist));

    if (ndist < ia->pdis) atrenergy = -ia->epsilon;
    else {
        atrenergy = cos(PIH * (ndist - ia->pdis) / ia->pswitch);
        atrenergy *= -atrenergy * ia->epsilon;
    }
    vec1 = perp_project(vec_intrs, p1Dir);
    a = dot(vec1, p1P->dir) / vsize(vec1);
    f1 = fangl_scale(a, ia->pcangl[0 + 2 * patchnum1], ia->pcanglsw[0 + 2 * patchnum1]);

    vec1 = scal(-1.0, vec_intrs);
    vec1 = perp_project(vec1, p2Dir);
    a = dot(vec1, p2P->dir) / vsize(vec1);
    f2 = fangl_scale(a, ia->pcangl[1 + 2 * patchnum2], ia->pcanglsw[1 + 2 * patchnum2]);

    paral = 1.0;
    if (ia->parallel != 0.0) paral = scparallel(ia->parallel, p1Dir, p2Dir);

    atrenergy *= f0 * f1 * f2 * paral;
    return atrenergy;
}

static inline int is_psc_family(int g) { return g == SCO_PSC || g == SCO_CHPSC || g == SCO_TPSC || g == SCO_TCHPSC; }
static inline int is_cpsc_family(int g) { return g == SCO_CPSC || g == SCO_CHCPSC || g == SCO_TCPSC || g == SCO_TCHCPSC; }
static inline int is_chiral(int g) { return g == SCO_CHPSC || g == SCO_CHCPSC || g == SCO_TCHPSC || g == SCO_TCHCPSC; }
static inline int is_two_patch(int g) { return g == SCO_TPSC || g == SCO_TCPSC || g == SCO_TCHPSC || g == SCO_TCHCPSC; }

/* functor kinds installed by PairE::initIntFCE, mc/paire.cpp:6-80 (later ifs overwrite earlier ones) */
enum { K_EBASIC = 0, K_SC_PSCCPSC, K_SC_CPSC, K_SC_PSC, K_SC_SCN, K_SC_SCA, K_SP_WCA, K_SP_COS2,
       K_MIX_SCASPA, K_MIX_PSCSPA, K_MIX_CPSCSPA };

static int functor_kind(int g, int o) {
    int k = K_EBASIC;
    if ((is_cpsc_family(g) && is_psc_family(o)) || (is_psc_family(g) && is_cpsc_family(o))) k = K_SC_PSCCPSC;
    if (is_cpsc_family(g) && is_cpsc_family(o)) k = K_SC_CPSC;
    if (is_psc_family(g) && is_psc_family(o)) k = K_SC_PSC;
    if (g == SCO_SCN && o == SCO_SCN) k = K_SC_SCN;
    if (g == SCO_SCA && o == SCO_SCA) k = K_SC_SCA;
    if (g == SCO_SPN || o == SCO_SPN) k = K_SP_WCA;
    if (g == SCO_SPA && o == SCO_SPA) k = K_SP_COS2;
    if ((g == SCO_SCA && o == SCO_SPA) || (g == SCO_SPA && o == SCO_SCA)) k = K_MIX_SCASPA;
    if ((is_psc_family(g) && (o == SCO_SPA || o == SCO_SPN)) || ((g == SCO_SPA || g == SCO_SPN) && is_psc_family(o))) k = K_MIX_PSCSPA;
    if ((is_cpsc_family(g) && (o == SCO_SPA || o == SCO_SPN)) || ((g == SCO_SPA || g == SCO_SPN) && is_cpsc_family(o))) k = K_MIX_CPSCSPA;
    return k;
}

/* Psc / CPsc / PscCPsc ::operator(), mc/paire.h:469-484, 658-672, 864-889 */
static double patch_e(int kind, const sco_iaparam* ia, vec3 p1Dir, vec3 p2Dir, const patch_t* p1P, const patch_t* p2P,
                      vec3 r_cm, int patchnum1, int patchnum2) {
    double T1, T2, S1, S2;
    vec3 vec1 = scal(-1.0, r_cm);
    int first_psc, second_psc;
    if (kind == K_SC_PSC) { first_psc = 1; second_psc = 1; }
    else if (kind == K_SC_CPSC) { first_psc = 0; second_psc = 0; }
    else { first_psc = is_psc_family((int)ia->geotype[0]); second_psc = !first_psc; }

    if (first_psc) {
        if (2 > psc_intersect(p1Dir, p2Dir, p1P, r_cm, &T1, &T2, ia->pcanglsw[2 * patchnum1], ia->rcutSq, ia->half_len[0], ia->half_len[1])) return 0.0;
    } else {
        if (2 > cpsc_intersect(p1Dir, p2Dir, p1P, r_cm, &T1, &T2, ia->pcanglsw[2 * patchnum1], ia->rcutSq, ia->half_len[0], ia->half_len[1])) return 0.0;
    }
    if (second_psc) {
        if (2 > psc_intersect(p2Dir, p1Dir, p2P, vec1, &S1, &S2, ia->pcanglsw[2 * patchnum2 + 1], ia->rcutSq, ia->half_len[1], ia->half_len[0])) return 0.0;
    } else {
        if (2 > cpsc_intersect(p2Dir, p1Dir, p2P, vec1, &S1, &S2, ia->pcanglsw[2 * patchnum2 + 1], ia->rcutSq, ia->half_len[1], ia->half_len[0])) return 0.0;
    }
    return atr_e(ia, p1Dir, p2Dir, p1P, p2P, r_cm, patchnum1, patchnum2, S1, S2, T1, T2);
}

/* EBond::harmonicPotential, mc/paire.h:202-204 */
static inline double harmonic(double aktualvalue, double eqvalue, double springconst) {
    return springconst * (aktualvalue - eqvalue) * (aktualvalue - eqvalue) * 0.5;
}

/* HarmonicSp::operator(), mc/paire.h:225-235 */
static double harmonic_sp(double dist, int i2, const sco_conlist* cl) {
    if (i2 == cl->con[1] || i2 == cl->con[0]) return harmonic(dist, cl->eq[0], cl->c[0]);
    if (i2 == cl->con[2] || i2 == cl->con[3]) return harmonic(dist, cl->eq[1], cl->c[1]);
    return 0.0;
}

/* HarmonicSc::operator(), mc/paire.h:244-279 */
static double harmonic_sc(const sco_system* s, double dist, const double* s1, const double* s2, const sco_iaparam* ia,
                          int i2, const sco_conlist* cl) {
    if (i2 == cl->con[1] || i2 == cl->con[0]) {
        double halfl1, halfl2, bondlength;
        int tail = (i2 == cl->con[0]);
        vec3 vec1, vec2;
        if ((int)ia->geotype[0] < SCO_SP) halfl1 = (ia->half_len[0] + cl->mod[tail ? 0 : 1]) * (tail ? 1.0 : -1.0);
        else halfl1 = cl->sp;
        if ((int)ia->geotype[1] < SCO_SP) halfl2 = (ia->half_len[1] + cl->mod[tail ? 1 : 0]) * (tail ? -1.0 : 1.0);
        else halfl2 = cl->sp;
        vec1 = V(s1[0] + (s1[3] * halfl1 / s->box[0]), s1[1] + (s1[4] * halfl1 / s->box[1]), s1[2] + (s1[5] * halfl1 / s->box[2]));
        vec2 = V(s2[0] + (s2[3] * halfl2 / s->box[0]), s2[1] + (s2[4] * halfl2 / s->box[1]), s2[2] + (s2[5] * halfl2 / s->box[2]));
        vec1 = image(s->box, vec1, vec2);
        bondlength = sqrt(dot(vec1, vec1));
        return harmonic(bondlength, cl->eq[0], cl->c[0]);
    }
    if (i2 == cl->con[2] || i2 == cl->con[3]) return harmonic(dist, cl->eq[1], cl->c[1]);
    return 0.0;
}

static inline vec3 normalised(vec3 v) { /* Vector::normalise, structures/Vector.h:56-64 */
    double tot = vsize(v);
    if (tot != 0.0) { tot = 1.0 / tot; v.x *= tot; v.y *= tot; v.z *= tot; }
    return v;
}

/* AngleSc::angleEnergyAngle2, mc/paire.h:340-352 */
static double angle2(const double* p1, const double* p2) {
    vec3 d1 = ld(p1 + 3), d2 = ld(p2 + 3);
    vec3 localAxis = cross(d1, d2);
    vec3 localX1 = cross(d1, localAxis);
    vec3 localX2 = cross(d2, localAxis);
    double v1x = dot(localX1, ld(p1 + 6)), v1y = dot(localAxis, ld(p1 + 6));
    double v2x = dot(localX2, ld(p2 + 6)), v2y = dot(localAxis, ld(p2 + 6));
    return acos((v1x * v2x + v1y * v2y) / (sqrt((v1x * v1x + v1y * v1y) * (v2x * v2x + v2y * v2y))));
}

/* AngleSc::operator(), mc/paire.h:287-337 */
static double angle_sc(const sco_system* s, const double* s1, int moltype1, const double* s2, const sco_iaparam* ia,
                       int i2, const sco_conlist* cl) {
    double energy = 0.0, currangle, halfl;
    const sco_molparam* mp = &s->mol[moltype1];
    int g0 = (int)ia->geotype[0], g1 = (int)ia->geotype[1];
    int near = (i2 == cl->con[0] || i2 == cl->con[1]);
    int tail = (i2 == cl->con[0]);
    if (mp->angle1c >= 0) {
        if (near) {
            vec3 vec1, vec2;
            if (g0 < SCO_SP) vec1 = ld(s1 + 3);
            else {
                halfl = ia->half_len[1] * (tail ? -1.0 : 1.0);
                vec1 = V(s2[0] + s2[3] * halfl / s->box[0], s2[1] + s2[4] * halfl / s->box[1], s2[2] + s2[5] * halfl / s->box[2]);
                vec1 = image(s->box, vec1, ld(s1));
                vec1 = normalised(vec1);
            }
            if (g1 < SCO_SP) vec2 = ld(s2 + 3);
            else {
                halfl = ia->half_len[0] * (tail ? 1.0 : -1.0);
                vec2 = V(s1[0] + s1[3] * halfl / s->box[0], s1[1] + s1[4] * halfl / s->box[1], s1[2] + s1[5] * halfl / s->box[2]);
                vec2 = image(s->box, vec2, ld(s2));
                vec2 = normalised(vec2);
            }
            currangle = acos(dot(vec1, vec2));
            energy += harmonic(currangle, mp->angle1eq, mp->angle1c);
        }
    }
    if (mp->angle2c >= 0) {
        if (near) {
            if ((g0 < SCO_SP) && (g1 < SCO_SP)) {
                currangle = tail ? angle2(s2, s1) : angle2(s1, s2);
                energy += harmonic(currangle, mp->angle2eq, mp->angle2c);
            }
        }
    }
    return energy;
}

/* WcaTruncSq, mc/paire.h:385-393 */
static inline double wca_trunc_sq(double distSq, const sco_iaparam* ia) {
    if (distSq > ia->rcutwcaSq) return 0.0;
    return ia->epsilon + ia->A * pow(distSq, -6) - ia->B * pow(distSq, -3);
}
/* WcaTrunc, mc/paire.h:376-383 */
static inline double wca_trunc(double dist, const sco_iaparam* ia) {
    if (dist > ia->rcutwca) return 0.0;
    return ia->A * pow(dist, -12) - ia->B * pow(dist, -6) + ia->epsilon;
}
/* WcaCos2Taylor, mc/paire.h:416-439 */
static double wca_cos2_taylor(double dist, const sco_iaparam* ia) {
    double e = 0.0;
    if (dist > ia->rcut || ia->epsilon == 0.0 || ia->exclude != 0.0) return 0.0;
    if (dist > ia->pdis) {
        e = PIH * (dist - ia->pdis) * ia->pswitchINV;
        e *= e;
        e = (1 - e + e * e * (1.0 / 3.0) - e * e * e * (2.0 / 45.0) + e * e * e * e * (1.0 / 315.0) - e * e * e * e * e * (2.0 / 14175.0)
             + e * e * e * e * e * e * (2.0 / 467775.0) - e * e * e * e * e * e * e * (4.0 / 42567525) + e * e * e * e * e * e * e * e * (1.0 / 638512875)) * -ia->epsilon;
    } else {
        e = -ia->epsilon;
    }
    if (dist > ia->rcutwca) return e;
    return ia->A * pow(dist, -12) - ia->B * pow(dist, -6);
}

/* SpheroCylinder<PatchE,HarmonicSc,AngleSc>::operator(), mc/paire.h:1122-1196 */
static double spherocylinder_e(const sco_system* s, int kind, double dist, vec3 r_cm, const double* s1, int moltype1,
                               const double* s2, const sco_iaparam* ia, int i2, const sco_conlist* cl) {
    double abE, distSq, atrenergy = 0.0, repenergy;
    vec3 dir1 = ld(s1 + 3), dir2 = ld(s2 + 3), dv;
    abE = harmonic_sc(s, dist, s1, s2, ia, i2, cl);
    abE += angle_sc(s, s1, moltype1, s2, ia, i2, cl);

    dv = min_dist_segments(dir1, dir2, ia->half_len[0], ia->half_len[1], r_cm);
    distSq = dot(dv, dv);
    repenergy = wca_trunc_sq(distSq, ia);

    if ((distSq > ia->rcutSq) || (ia->epsilon == 0.0) || ia->exclude != 0.0) {
        atrenergy = 0.0;
    } else if (kind == K_SC_SCN) {
        atrenergy = 0.0; /* Scn inherits EPatch::operator() which returns 0, paire.h:58-61, 829-840 */
    } else if (kind == K_SC_SCA) { /* Sca::operator(), paire.h:845-852 */
        double d = sqrt(distSq);
        if (d > ia->rcutwca) atrenergy = 0.0;
        else atrenergy = ia->A * pow(d, -12) - ia->B * pow(d, -6) + ia->epsilon;
    } else {
        int g0 = (int)ia->geotype[0], g1 = (int)ia->geotype[1];
        int firstCH = is_chiral(g0), secondCH = is_chiral(g1), firstT = is_two_patch(g0), secondT = is_two_patch(g1);
        patch_t P1a, P1b, P2a, P2b;
        vec3 ax1a = firstCH ? ld(s1 + 24) : dir1, ax1b = firstCH ? ld(s1 + 27) : dir1;
        vec3 ax2a = secondCH ? ld(s2 + 24) : dir2, ax2b = secondCH ? ld(s2 + 27) : dir2;
        P1a.dir = ld(s1 + 6); P1a.sides[0] = ld(s1 + 12); P1a.sides[1] = ld(s1 + 15);
        P1b.dir = ld(s1 + 9); P1b.sides[0] = ld(s1 + 18); P1b.sides[1] = ld(s1 + 21);
        P2a.dir = ld(s2 + 6); P2a.sides[0] = ld(s2 + 12); P2a.sides[1] = ld(s2 + 15);
        P2b.dir = ld(s2 + 9); P2b.sides[0] = ld(s2 + 18); P2b.sides[1] = ld(s2 + 21);
        atrenergy = patch_e(kind, ia, ax1a, ax2a, &P1a, &P2a, r_cm, 0, 0);
        if (firstT) atrenergy += patch_e(kind, ia, ax1b, ax2a, &P1b, &P2a, r_cm, 1, 0);
        if (secondT) atrenergy += patch_e(kind, ia, ax1a, ax2b, &P1a, &P2b, r_cm, 0, 1);
        if (firstT && secondT) atrenergy += patch_e(kind, ia, ax1b, ax2b, &P1b, &P2b, r_cm, 1, 1);
    }
    return abE + repenergy + atrenergy;
}

/* MixSpSc::closestDist, mc/paire.h:1077-1093 */
static inline double closest_dist_sp(vec3 r_cm, vec3 dir1, double halfl, double* contt, vec3* distvec) {
    double d;
    *contt = dot(dir1, r_cm);
    d = -halfl;
    if (*contt >= halfl) d = halfl;
    else if (*contt > -halfl) d = *contt;
    distvec->x = -r_cm.x + dir1.x * d;
    distvec->y = -r_cm.y + dir1.y * d;
    distvec->z = -r_cm.z + dir1.z * d;
    return dot(*distvec, *distvec);
}

/* PscSpa / ScaSpa ::operator(), mc/paire.h:914-979 */
static double patch_to_sphere(int kind, double dist, double contt, vec3 distvec, const sco_iaparam* ia, vec3 p1Dir, vec3 patchdir) {
    double atrenergy, a, b, f0, halfl;
    if (dist < ia->pdis) atrenergy = -ia->epsilon;
    else {
        atrenergy = cos(PIH * (dist - ia->pdis) / ia->pswitch);
        atrenergy *= -atrenergy * ia->epsilon;
    }
    halfl = ia->half_len[0];
    b = sqrt(ia->rcutSq - dist * dist);
    if (contt + b > halfl) f0 = halfl; else f0 = contt + b;
    if (contt - b < -halfl) f0 -= -halfl; else f0 -= contt - b;
    if (kind == K_MIX_SCASPA) {
        atrenergy *= f0;
    } else {
        vec3 vec1 = perp_project(distvec, p1Dir);
        a = dot(vec1, patchdir) / vsize(vec1);
        atrenergy *= fangl_scale(a, ia->pcangl[0], ia->pcanglsw[0]) * (f0);
    }
    return atrenergy;
}

/* MixSpSc<PatchE,HarmonicSc,AngleSc>::operator(), mc/paire.h:1022-1067 */
static double mix_sp_sc_e(const sco_system* s, int kind, double dist, vec3 r_cm, const double* s1, int type1, int moltype1,
                          const double* s2, int type2, int i2, const sco_conlist* cl) {
    const sco_iaparam* ia12 = &s->ia[type1 * s->ntypes + type2];
    int isp1Spc = ((int)ia12->geotype[0] < SCO_SP);
    const double* spc = isp1Spc ? s1 : s2;
    const sco_iaparam* ia = isp1Spc ? ia12 : &s->ia[type2 * s->ntypes + type1];
    vec3 rr = isp1Spc ? r_cm : scal(-1.0, r_cm);
    double atrenergy = 0.0, repenergy = 0.0, contt = 0.0, abE;
    vec3 distvec;
    double distSq = closest_dist_sp(rr, ld(spc + 3), ia->half_len[0], &contt, &distvec);

    abE = harmonic_sc(s, dist, s1, s2, ia12, i2, cl);
    abE += angle_sc(s, s1, moltype1, s2, ia12, i2, cl);

    if (distSq < ia->rcutwcaSq) repenergy = wca_trunc_sq(distSq, ia12);

    {
        int g0 = (int)ia->geotype[0];
        int chiral = 0, sec = 0, is_far = 0;
        if (kind == K_MIX_PSCSPA) { chiral = (g0 == SCO_CHPSC || g0 == SCO_TCHPSC); sec = (g0 == SCO_TPSC || g0 == SCO_TCHPSC); }
        if (kind == K_MIX_CPSCSPA) {
            chiral = (g0 == SCO_CHCPSC || g0 == SCO_TCHCPSC); sec = (g0 == SCO_TCPSC || g0 == SCO_TCHCPSC);
            is_far = (contt > ia->half_len[0]) || (contt < -ia->half_len[0]); /* CPscSpa::isFar, paire.h:1007-1009 */
        }
        if ((distSq > ia->rcutSq) || (ia->epsilon == 0.0) || ia->exclude != 0.0 || is_far) {
            atrenergy = 0.0;
        } else {
            vec3 ax0 = chiral ? ld(spc + 24) : ld(spc + 3);
            vec3 ax1 = chiral ? ld(spc + 27) : ld(spc + 3);
            if (chiral) distSq = closest_dist_sp(rr, ax0, ia->half_len[0], &contt, &distvec);
            dist = sqrt(distSq);
            if (dist < ia->rcut) atrenergy = patch_to_sphere(kind, dist, contt, distvec, ia, ax0, ld(spc + 6));
            if (sec) {
                distSq = closest_dist_sp(rr, ax1, ia->half_len[0], &contt, &distvec);
                dist = sqrt(distSq);
                if (dist < ia->rcut) atrenergy += patch_to_sphere(kind, dist, contt, distvec, ia, ax1, ld(spc + 9));
            }
        }
    }
    return abE + repenergy + atrenergy;
}


/* ------------------------------------------------------------------------------------------------
 * External wall potential ([EXTER]): ExternalEnergyCalculator::extere2 and its helpers, scOOP/mc/externalenergycalculator.cpp:5-500
 * (externalenergycalculator.h:21-115); parameters topo.exter.interactions[], scOOP/structures/topo.cpp:120-130, 151-152.
 * The reference keeps intermediate values in members of the calculator object; they are the fields of wall_state here.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { double sigma, epsilon, rcutwca, rcut, pdis, pswitch, len0, half_len0; int geotype, pad; } wall_param;
typedef struct { double dist, orient, rcmz, interendz; int positive, orientin; vec3 project; vec3 dir; } wall_state;
#define false 0
#define true 1
static vec3 wall_v3zero(void) { vec3 v; v.x = 0; v.y = 0; v.z = 0; return v; }
#define wall_project_in_z(v, pd, out) wall_project_in_z_((v), (pd), &(out))
static void wall_project_in_z_(vec3 vec1, vec3 projectdir, vec3* projection) {
    projection->x = vec1.x - vec1.z * projectdir.x / projectdir.z;
    projection->y = vec1.y - vec1.z * projectdir.y / projectdir.z;
    projection->z = 0;
}

/* ExternalEnergyCalculator::exter2ClosestDist (:107-145) */
static void wall_closest_dist(wall_state* w, const wall_param* param) {
    if (w->rcmz < 0) { w->dist = -(w->rcmz); w->positive = false; w->interendz = -1.0; w->project.z = 1.0; }
    else { w->dist = (w->rcmz); w->positive = true; w->interendz = 1.0; w->project.z = -1.0; }
    if (param->geotype < SCO_SPN) {        /* psc closest is always the end closer to the wall */
        if (w->dir.z > 0) {
            if (w->positive) { w->orientin = false; w->orient = -1.0; w->dist = w->rcmz - w->dir.z * param->half_len0; }
            else { w->orientin = true; w->orient = 1.0; w->dist = -(w->rcmz + w->dir.z * param->half_len0); }
        } else {
            if (w->positive) { w->orientin = true; w->orient = 1.0; w->dist = w->rcmz + w->dir.z * param->half_len0; }
            else { w->orientin = false; w->orient = -1.0; w->dist = -(w->rcmz - w->dir.z * param->half_len0); }
        }
    }
}

/* ExternalEnergyCalculator::pscWall (:263-458) */
#define pbeg (*pbeg_)
#define pend (*pend_)
static int wall_psc_(const wall_state* w, vec3* pbeg_, vec3* pend_, vec3 projectdir, vec3 partdir, double cutdist,
                                        vec3 partbeg, vec3 partend) {
    vec3 vec1;
    double k, x1, x2, y1, y2, a, b, c, e, d;
    if (((w->positive) && (projectdir.z > 0)) || ((!(w->positive)) && (projectdir.z < 0))) return 0;
    if (fabs(partbeg.z) > cutdist) return 0;
    x2 = 0.0;
    y2 = 0.0;
    if (fabs(partdir.z) > 1.0e-8) { wall_project_in_z(partbeg, partdir, pbeg); a = 0; }
    else {
        vec1.x = 2.0 * partbeg.x - partend.x;
        vec1.y = 2.0 * partbeg.y - partend.y;
        vec1.z = 2.0 * partbeg.z - partend.z;
        wall_project_in_z(vec1, projectdir, pbeg);
        a = 1;
    }
    if (partdir.z != 0) b = fabs(partbeg.z / partdir.z);
    else b = cutdist + 1.0;
    if ((b > cutdist) || (a == 1)) {
        if (fabs(projectdir.z) > 1.0e-8) wall_project_in_z(partbeg, projectdir, pend);
        else { pend.x = pbeg.x + projectdir.x; pend.y = pbeg.y + projectdir.y; }
        if (pend.y == pbeg.y) {
            y1 = pbeg.y;
            y2 = pbeg.y;
            a = sqrt(cutdist * cutdist - partbeg.z * partbeg.z - (pbeg.y - partbeg.y) * (pbeg.y - partbeg.y));
            x1 = partbeg.x + a;
            x2 = partbeg.x - a;
            if (pend.x > pbeg.x) { pbeg.x = x2; x2 = x1; }
            else pbeg.x = x1;
            pbeg.y = y1;
        } else {
            k = (pend.x - pbeg.x) / (pend.y - pbeg.y);
            a = k * k + 1;
            b = partbeg.y + k * k * pbeg.y - k * pbeg.x + k * partbeg.x;
            c = partbeg.y * partbeg.y + partbeg.z * partbeg.z - cutdist * cutdist + (k * pbeg.y - pbeg.x + partbeg.x) * (k * pbeg.y - pbeg.x + partbeg.x);
            e = b * b - a * c;
            if (e < 0) return 0;
            d = sqrt(e);
            if (pend.y > pbeg.y) { y1 = (b - d) / a; y2 = (b + d) / a; }
            else { y1 = (b + d) / a; y2 = (b - d) / a; }
            x1 = k * (y1 - pbeg.y) + pbeg.x;
            x2 = k * (y2 - pbeg.y) + pbeg.x;
            pbeg.x = x1;
            pbeg.y = y1;
            pbeg.z = 0.0;
        }
    }
    /* end point */
    a = -cutdist * projectdir.z;      /* z coordinate of the point where the projection is at the cut distance */
    if (((partend.z < a) && (w->positive)) || ((a < partend.z) && (!(w->positive)))) {
        if (projectdir.z != 0) wall_project_in_z(partend, projectdir, pend);
        else { pend.x = pbeg.x + projectdir.x; pend.y = pbeg.y + projectdir.y; }
        if (pend.y == pbeg.y) {
            y1 = pend.y;
            y2 = pend.y;
            a = sqrt(cutdist * cutdist - partend.z * partend.z - (pend.y - partend.y) * (pend.y - partend.y));
            x1 = partend.x + a;
            x2 = partend.x - a;
            if (pbeg.x > pend.x) pend.x = x2;
            else pend.x = x1;
            pend.y = y1;
        } else {
            k = (pbeg.x - pend.x) / (pbeg.y - pend.y);
            a = k * k + 1;
            b = partend.y + k * k * pend.y - k * pend.x + k * partend.x;
            c = partend.y * partend.y + partend.z * partend.z - cutdist * cutdist + (k * pend.y - pend.x + partend.x) * (k * pend.y - pend.x + partend.x);
            e = b * b - a * c;
            if (e < 0) return 0;
            d = sqrt(e);
            if (pbeg.y > pend.y) { y1 = (b - d) / a; y2 = (b + d) / a; }
            else { y1 = (b + d) / a; y2 = (b - d) / a; }
            x1 = k * (y1 - pend.y) + pend.x;
            x2 = k * (y2 - pend.y) + pend.x;
            pend.x = x1;
            pend.y = y1;
            pend.z = 0.0;
        }
    } else {
        if (((partbeg.z < a) && (w->positive)) || ((a < partbeg.z) && (!(w->positive)))) {
            /* the end is at the cutoff, going through the cylindrical part */
            b = (a - partbeg.z) / partdir.z;
            vec1.x = partbeg.x + b * partdir.x;
            vec1.y = partbeg.y + b * partdir.y;
            vec1.z = a;
            wall_project_in_z(vec1, projectdir, pend);
        } else {
            /* the projected end is within the same sphere as the beginning: no contribution from the cylinder */
            if (x2 == 0.0) {
                if (projectdir.z != 0) wall_project_in_z(partbeg, projectdir, pend);
                else { pend.x = pbeg.x + projectdir.x; pend.y = pbeg.y + projectdir.y; }
                if (pend.y == pbeg.y) {
                    y1 = pbeg.y;
                    y2 = pbeg.y;
                    a = sqrt(cutdist * cutdist - partbeg.z * partbeg.z - (pbeg.y - partbeg.y) * (pbeg.y - partbeg.y));
                    x1 = partbeg.x + a;
                    x2 = partbeg.x - a;
                    if (pend.x > pbeg.x) pend.x = x1;
                    else pend.x = x2;
                    pend.y = y1;
                } else {
                    k = (pend.x - pbeg.x) / (pend.y - pbeg.y);
                    a = k * k + 1;
                    b = partbeg.y + k * k * pbeg.y - k * pbeg.x + k * partbeg.x;
                    c = partbeg.y * partbeg.y + partbeg.z * partbeg.z - cutdist * cutdist + (k * pbeg.y - pbeg.x + partbeg.x) * (k * pbeg.y - pbeg.x + partbeg.x);
                    e = b * b - a * c;
                    if (e < 0) return 0;
                    d = sqrt(e);
                    if (pend.y > pbeg.y) { y1 = (b - d) / a; y2 = (b + d) / a; }
                    else { y1 = (b + d) / a; y2 = (b - d) / a; }
                    x1 = k * (y1 - pbeg.y) + pbeg.x;
                    x2 = k * (y2 - pbeg.y) + pbeg.x;
                    pend.x = x1;
                    pend.y = y1;
                    pend.z = 0.0;
                }
            } else {
                pend.x = x2;
                pend.y = y2;
                pend.z = 0.0;
            }
            return 2;
        }
    }
    return 1;
}

/* ExternalEnergyCalculator::cpscWall (:460-500) */
static int wall_cpsc_(const wall_state* w, vec3* pbeg_, vec3* pend_, vec3 projectdir, vec3 partdir, double halfl,
                                         double cutdist, vec3 partbeg, vec3 partend) {
    vec3 vec1;
    double a;
    if (((w->positive) && (projectdir.z >= 0)) || ((!(w->positive)) && (projectdir.z <= 0))) return 0;
    vec1.x = partbeg.x;
    vec1.y = partbeg.y;
    vec1.z = partbeg.z;
    if (-vec1.z / projectdir.z < cutdist) wall_project_in_z(vec1, projectdir, pbeg);
    else return 0;
    if (-partend.z / projectdir.z < cutdist) vec1.z = partend.z;
    else vec1.z = -cutdist * projectdir.z;
    if (partdir.z != 0.0) a = (vec1.z - (w->rcmz)) / partdir.z;
    else { if (w->orientin) a = -halfl; else a = halfl; }
    vec1.x = partdir.x * a;
    vec1.y = partdir.y * a;
    wall_project_in_z(vec1, projectdir, pend);
    return 1;
}

#undef pbeg
#undef pend
#define wall_psc(w, pb, pe, a, b, c, d, e) wall_psc_((w), &(pb), &(pe), (a), (b), (c), (d), (e))
#define wall_cpsc(w, pb, pe, a, b, c, d, e, f) wall_cpsc_((w), &(pb), &(pe), (a), (b), (c), (d), (e), (f))
/* ExternalEnergyCalculator::exter2Atre (:147-261) */
static double wall_atre(wall_state* w, const wall_param* param, double* ndist, vec3 patchdir, double halfl) {
    vec3 pbeg, pend;
    double a, length1, length2, f0, f1;
    vec3 cm1, cm2;
    int line;
    vec3 partbeg, partend;
    vec3 inters;
    double atrenergy = 0.0;
    pbeg = wall_v3zero(); pend = wall_v3zero();
    if ((param->geotype < SCO_SPN) && (param->geotype > SCO_SCA)) {
        a = ((w->orientin ? 1.0 : 0.0) - 0.5) * 2;
        partbeg.x = a * w->dir.x * halfl;
        partbeg.y = a * w->dir.y * halfl;
        partbeg.z = w->rcmz + a * w->dir.z * halfl;
        partend.x = -a * w->dir.x * halfl;
        partend.y = -a * w->dir.y * halfl;
        partend.z = w->rcmz - a * w->dir.z * halfl;
        if ((param->rcut - w->dist) / fabs(w->dir.z) < 2.0 * halfl) w->interendz *= param->rcut;
        else w->interendz = partend.z;
        if (w->positive) cm1.z = ((w->interendz + w->dist) * 0.5);
        else cm1.z = ((w->interendz + -w->dist) * 0.5);
        if (w->dir.z != 0.0) {
            a = (w->interendz - cm1.z) / w->dir.z;
            length1 = -w->orient * 2.0 * a;
            a = a + w->orient * halfl;
        } else {
            a = 0.0;
            length1 = 2.0 * halfl;
        }
        cm1.x = w->dir.x * a;
        cm1.y = w->dir.y * a;
        if ((param->geotype == SCO_CPSC) || (param->geotype == SCO_CHCPSC)) {
            if (((w->interendz >= w->dist) && (w->positive)) || ((w->interendz <= -w->dist) && (!(w->positive))))
                line = wall_cpsc(w, pbeg, pend, w->project, w->dir, param->half_len0, param->rcut, partbeg, partend);
            else line = 0;
        } else {
            line = wall_psc(w, pbeg, pend, w->project, w->dir, param->rcut, partbeg, partend);
        }
        if (line > 0) {
            cm2.x = ((pbeg.x + pend.x) * 0.5);
            cm2.y = ((pbeg.y + pend.y) * 0.5);
            cm2.z = 0.0;
            length2 = sqrt((pend.x - pbeg.x) * (pend.x - pbeg.x) + (pend.y - pbeg.y) * (pend.y - pbeg.y));
            inters.x = cm2.x - cm1.x;
            inters.y = cm2.y - cm1.y;
            inters.z = cm2.z - cm1.z;
            *ndist = sqrt(inters.x * inters.x + inters.y * inters.y + inters.z * inters.z);
            if (*ndist < param->pdis) atrenergy = -param->epsilon;
            else {
                atrenergy = cos(PIH * (*ndist - param->pdis) / param->pswitch);
                atrenergy *= -atrenergy * param->epsilon;
            }
            f0 = (length1 + length2) * 0.5;
            f1 = fabs(patchdir.z);
            atrenergy *= f0 * f1;
        } else atrenergy = 0.0;
    } else {
        if (*ndist < param->pdis) atrenergy = -param->epsilon;
        else {
            atrenergy = cos(PIH * (*ndist - param->pdis) / param->pswitch);
            atrenergy *= -atrenergy * param->epsilon;
        }
        atrenergy *= (param->rcut * param->rcut - (*ndist) * (*ndist)) / (param->sigma * param->sigma);
    }
    return atrenergy;
}

/* ExternalEnergyCalculator::extere2 (:5-105). pos_z: box-fractional z; dir, patchdir[2], chdir[2]: the particle's vectors. */
static double wall_energy(const wall_param* param, double exter_sqmaxcut, double box_z, double pos_z, vec3 dir,
                                              vec3 patchdir0, vec3 patchdir1, vec3 chdir0, vec3 chdir1) {
    double repenergy = 0.0, atrenergy = 0.0;
    double ndist, halfl;
    wall_state w_;
    wall_state* w = &w_;
    if (pos_z < 0) w->rcmz = box_z * (pos_z - (double)((long long)(pos_z - 0.5)));
    else w->rcmz = box_z * (pos_z - (double)((long long)(pos_z + 0.5)));
    w->project = wall_v3zero();
    if (w->rcmz < 0) { w->dist = -w->rcmz; w->positive = false; w->interendz = -1.0; w->project.z = 1.0; }
    else { w->dist = w->rcmz; w->positive = true; w->interendz = 1.0; w->project.z = -1.0; }
    if (w->rcmz * w->rcmz > exter_sqmaxcut) return 0.0;
    halfl = 0.5 * param->len0;
    ndist = w->dist;
    w->orientin = true;
    w->orient = 0.0;
    w->dir = dir;
    wall_closest_dist(w, param);
    if (w->dist > param->rcutwca) repenergy = 0.0;
    else {
        const double en6 = pow((param->sigma / w->dist), 6);
        repenergy = 4 * en6 * (en6 - 1) + 1.0;
    }
    if ((param->geotype == SCO_CHCPSC) || (param->geotype == SCO_CHPSC)) {
        w->dir = chdir0;
        wall_closest_dist(w, param);
    }
    if ((w->dist > param->rcut) || (param->epsilon == 0.0) || ((patchdir0.z > 0) && (w->positive)) || ((patchdir0.z < 0) && (!w->positive))) atrenergy = 0.0;
    else atrenergy = wall_atre(w, param, &ndist, patchdir0, halfl);
    if ((param->geotype == SCO_TCPSC) || (param->geotype == SCO_TPSC) || (param->geotype == SCO_TCHCPSC) || (param->geotype == SCO_TCHPSC)) {
        if ((param->geotype == SCO_TCHCPSC) || (param->geotype == SCO_TCHPSC)) {
            w->dir = chdir1;
            wall_closest_dist(w, param);
        }
        wall_closest_dist(w, param);
        if ((w->dist > param->rcut) || (param->epsilon == 0.0) || ((patchdir1.z > 0) && (w->positive)) || ((patchdir1.z < 0) && (!(w->positive)))) atrenergy += 0.0;
        else atrenergy += wall_atre(w, param, &ndist, patchdir1, halfl);
    }
    return repenergy + atrenergy;
}


#undef wall_psc
#undef wall_cpsc
#undef wall_project_in_z
#undef false
#undef true

/* topo.exter.interactions[type] out of the type's own table entry (structures/topo.cpp:120-130); out8 = sigma, epsilon, rcutwca, rcut,
 * pdis, pswitch, len0, half_len0 */
void sco_exter_params(const sco_iaparam* self, double thickness, double epsilon, double attraction, double* out8) {
    double sigma = (self->sigma + thickness) * 0.5;
    double rcutwca = (sigma) * pow(2.0, 1.0 / 6.0);
    double eps = sqrt(self->epsilon * epsilon);
    double pswitch = (self->pswitch + attraction) * 0.5;
    double pdis = (self->pdis - self->rcutwca + 0.0) * 0.5 + rcutwca;
    out8[0] = sigma; out8[1] = eps; out8[2] = rcutwca; out8[3] = pswitch + pdis; out8[4] = pdis; out8[5] = pswitch;
    out8[6] = self->len[0]; out8[7] = self->half_len[0];
}

/* extere2 of one particle (30-double state record); p8 as above */
double sco_extere2(const double* st, int geotype, const double* p8, double exter_sqmaxcut, double box_z) {
    wall_param pr;
    pr.sigma = p8[0]; pr.epsilon = p8[1]; pr.rcutwca = p8[2]; pr.rcut = p8[3]; pr.pdis = p8[4]; pr.pswitch = p8[5];
    pr.len0 = p8[6]; pr.half_len0 = p8[7]; pr.geotype = geotype; pr.pad = 0;
    return wall_energy(&pr, exter_sqmaxcut, box_z, st[2], ld(st + 3), ld(st + 6), ld(st + 9), ld(st + 24), ld(st + 27));
}

/* Hooks of the op-counting build (oracle/flopcount.cpp); they expand to nothing in the oracle proper. */
#ifndef SCO_COUNT_ENTER
#define SCO_COUNT_ENTER()
#define SCO_COUNT_GATED()
#define SCO_COUNT_PAUSE()
#define SCO_COUNT_RESUME()
#endif

/* PairE::operator(), mc/paire.h:1209-1220 */
double sco_pair_energy(const sco_system* s, const double* s1, int type1, int moltype1, int i1,
                       const double* s2, int type2, int i2, const sco_conlist* cl) {
    vec3 r_cm;
    double dotrcm;
    const sco_iaparam* ia = &s->ia[type1 * s->ntypes + type2];
    double dist;
    int kind;
    (void)i1;
    SCO_COUNT_ENTER();
    r_cm = image(s->box, ld(s1), ld(s2));
    dotrcm = dot(r_cm, r_cm);
    SCO_COUNT_GATED();          /* everything up to here is the cutoff gate: executed for EVERY candidate */
    if (dotrcm > s->sqmaxcut && cl->is_empty) return 0.0;
    dist = sqrt(dotrcm);
    kind = functor_kind((int)ia->geotype[0], (int)ia->geotype[1]);
    switch (kind) {
    case K_SC_PSCCPSC: case K_SC_CPSC: case K_SC_PSC: case K_SC_SCN: case K_SC_SCA:
        return spherocylinder_e(s, kind, dist, r_cm, s1, moltype1, s2, ia, i2, cl);
    case K_SP_WCA: /* Sphere<WcaTrunc,HarmonicSp>, paire.h:1106-1108 */
        return harmonic_sp(dist, i2, cl) + wca_trunc(dist, ia);
    case K_SP_COS2:
        return harmonic_sp(dist, i2, cl) + wca_cos2_taylor(dist, ia);
    case K_MIX_SCASPA: case K_MIX_PSCSPA: case K_MIX_CPSCSPA:
        return mix_sp_sc_e(s, kind, dist, r_cm, s1, type1, moltype1, s2, type2, i2, cl);
    default: /* EBasic: "not programmed", returns 0, paire.h:209-212 */
        return 0.0;
    }
}

/* ParticleVector::getConlist, structures/Conf.h:90-147 */
void sco_get_conlist(const sco_system* s, int i, sco_conlist* cl) {
    const sco_molparam* mp = &s->mol[s->moltype[i]];
    int msize = (int)mp->mol_size, first = (int)mp->first, pos;
    cl->is_empty = 1;
    cl->con[0] = cl->con[1] = cl->con[2] = cl->con[3] = -1;
    cl->sp = 0.0; cl->mod[0] = cl->mod[1] = 0.0; cl->c[0] = cl->c[1] = 0.0; cl->eq[0] = cl->eq[1] = 0.0;
    if (msize == 1) return;
    pos = (i - first) % msize;
    if (mp->bond1c >= 0.0 || mp->bonddc >= 0.0 || mp->bondhc >= 0.0) {
        if (pos > 0) cl->con[0] = i - 1;
        if (pos + 1 < msize) cl->con[1] = i + 1;
        if (mp->bond1c >= 0.0) { cl->eq[0] = mp->bond1eq; cl->c[0] = mp->bond1c; cl->mod[0] = 0.0; cl->mod[1] = 0.0; cl->sp = 0.0; }
        if (mp->bonddc >= 0.0) { cl->eq[0] = 0.0; cl->c[0] = mp->bonddc; cl->mod[0] = mp->bonddeq; cl->mod[1] = 0.0; cl->sp = 0.0; }
        if (mp->bondhc >= 0.0) { cl->eq[0] = 0.0; cl->c[0] = mp->bondhc; cl->mod[0] = mp->bondheq; cl->mod[1] = mp->bondheq; cl->sp = mp->bondheq; }
        cl->is_empty = 0;
    }
    if (mp->bond2c >= 0.0) {
        if (pos > 1) cl->con[2] = i - 2;
        if (pos + 2 < msize) cl->con[3] = i + 2;
        cl->eq[1] = mp->bond2eq; cl->c[1] = mp->bond2c;
        cl->is_empty = 0;
    }
}

/* TotalEFull::oneToAll, mc/totalenergycalculator.h:563-583 (== TotalEMatrix::oneToAllTrial :383-415) */
double sco_one_to_all(const sco_system* s, int target, const double* trial_state, double* e_pairs) {
    sco_conlist cl;
    const double* st = trial_state ? trial_state : s->state + (size_t)target * SCO_STATE;
    double energy = 0.0;
    int i;
    sco_get_conlist(s, target, &cl);
    for (i = 0; i < s->n; i++) {
        double e = 0.0;
        if (i != target) {
            e = sco_pair_energy(s, st, s->type[target], s->moltype[target], target,
                                s->state + (size_t)i * SCO_STATE, s->type[i], i, &cl);
            energy += e;
        }
        if (e_pairs) e_pairs[i] = e;
    }
    return energy;
}

/* TotalEMatrix::allToAll(matrix), mc/totalenergycalculator.h:502-521 */
double sco_all_to_all(const sco_system* s, double* e_row) {
    double energy = 0.0;
    int i, j;
    if (e_row && s->n > 0) e_row[0] = 0.0;
    for (i = 1; i < s->n; i++) {
        sco_conlist cl;
        double row = 0.0;
        sco_get_conlist(s, i, &cl);
        for (j = 0; j < i; j++) {
            double e = sco_pair_energy(s, s->state + (size_t)i * SCO_STATE, s->type[i], s->moltype[i], i,
                                       s->state + (size_t)j * SCO_STATE, s->type[j], j, &cl);
            energy += e;
            row += e;
        }
        if (e_row) e_row[i] = row;
    }
    return energy;
}

/* TotalEMatrix::mol2othersTrial (no pair list), mc/totalenergycalculator.h:455-499 */
double sco_mol_to_others(const sco_system* s, int first, int m) {
    sco_conlist empty;
    double energy = 0.0;
    int j, i;
    memset(&empty, 0, sizeof(empty));
    empty.is_empty = 1;
    empty.con[0] = empty.con[1] = empty.con[2] = empty.con[3] = -1;
    for (j = first; j < first + m; j++) {
        for (i = 0; i < first; i++)
            energy += sco_pair_energy(s, s->state + (size_t)j * SCO_STATE, s->type[j], s->moltype[j], j,
                                      s->state + (size_t)i * SCO_STATE, s->type[i], i, &empty);
        for (i = first + m; i < s->n; i++)
            energy += sco_pair_energy(s, s->state + (size_t)j * SCO_STATE, s->type[j], s->moltype[j], j,
                                      s->state + (size_t)i * SCO_STATE, s->type[i], i, &empty);
    }
    return energy;
}

/* Vector::rotate(axis, cos, sin), structures/Vector.h:138-160 */
static vec3 rotate_q(vec3 p, vec3 axis, double cosAngle, double sinAngle) {
    double t2, t3, t4, t5, t6, t7, t8, t9, t10;
    double qw = cosAngle, qx = (axis.x * sinAngle), qy = (axis.y * sinAngle), qz = (axis.z * sinAngle);
    vec3 r;
    t2 = qw * qx; t3 = qw * qy; t4 = qw * qz;
    t5 = -qx * qx; t6 = qx * qy; t7 = qx * qz;
    t8 = -qy * qy; t9 = qy * qz; t10 = -qz * qz;
    r.x = 2.0 * ((t8 + t10) * p.x + (t6 - t4) * p.y + (t3 + t7) * p.z) + p.x;
    r.y = 2.0 * ((t4 + t6) * p.x + (t5 + t10) * p.y + (t9 - t2) * p.z) + p.y;
    r.z = 2.0 * ((t7 - t3) * p.x + (t2 + t9) * p.y + (t5 + t8) * p.z) + p.z;
    return r;
}
static inline void st3(double* p, vec3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
/* Vector::ortogonalise, structures/Vector.h:133-135 */
static inline vec3 ortogonalise(vec3 a, vec3 B) {
    double dp = dot(a, B);
    a.x -= dp * B.x; a.y -= dp * B.y; a.z -= dp * B.z;
    return a;
}

/* Particle::init, structures/particle.cpp:3-79 */
void sco_particle_init(const sco_iaparam* ia, double* st) {
    int g = (int)ia->geotype[0];
    vec3 dir, pd0, pd1, ch0, ch1;
    if (g == SCO_SCA || g == SCO_SCN) return;
    dir = normalised(ld(st + 3));
    pd0 = normalised(ortogonalise(ld(st + 6), dir));
    st3(st + 3, dir); st3(st + 6, pd0);
    pd1 = ld(st + 9);
    if (g == SCO_PSC || g == SCO_CPSC || g == SCO_TPSC || g == SCO_TCPSC) {
        st3(st + 12, rotate_q(pd0, dir, ia->pcoshalfi[0], ia->psinhalfi[0]));
        st3(st + 15, rotate_q(pd0, dir, ia->pcoshalfi[0], -1.0 * ia->psinhalfi[0]));
    }
    if (is_two_patch(g)) {
        pd1 = rotate_q(pd0, dir, ia->csecpatchrot[0], ia->ssecpatchrot[0]);
        pd1 = normalised(ortogonalise(pd1, dir));
        st3(st + 9, pd1);
    }
    if (g == SCO_TPSC || g == SCO_TCPSC) {
        st3(st + 18, rotate_q(pd1, dir, ia->pcoshalfi[2], ia->psinhalfi[2]));
        st3(st + 21, rotate_q(pd1, dir, ia->pcoshalfi[2], -1.0 * ia->psinhalfi[2]));
    }
    if (is_chiral(g)) {
        ch0 = rotate_q(dir, pd0, ia->chiral_cos[0], ia->chiral_sin[0]);
        st3(st + 24, ch0);
        st3(st + 12, rotate_q(pd0, ch0, ia->pcoshalfi[0], ia->psinhalfi[0]));
        st3(st + 15, rotate_q(pd0, ch0, ia->pcoshalfi[0], -1.0 * ia->psinhalfi[0]));
    }
    if (g == SCO_TCHPSC || g == SCO_TCHCPSC) {
        ch1 = rotate_q(dir, pd1, ia->chiral_cos[0], ia->chiral_sin[0]);
        st3(st + 27, ch1);
        st3(st + 18, rotate_q(pd1, ch1, ia->pcoshalfi[2], ia->psinhalfi[2]));
        st3(st + 21, rotate_q(pd1, ch1, ia->pcoshalfi[2], -1.0 * ia->psinhalfi[2]));
    }
}

/* Particle::pscRotate, structures/particle.h:182-272; the random sense is passed in */
void sco_psc_rotate(double* st, int geotype, double angle, const double axis[3], int clockwise_positive) {
    double vc = cos(angle), vs;
    double qw, qx, qy, qz, t2, t3, t4, t5, t6, t7, t8, t9, t10, d1, d2, d3, d4, d5, d6, d7, d8, d9;
    int k, m, nv = 0, idx[12];
    if (clockwise_positive) vs = sqrt(1.0 - vc * vc); else vs = -sqrt(1.0 - vc * vc);
    qw = vc; qx = axis[0] * vs; qy = axis[1] * vs; qz = axis[2] * vs;
    t2 = qw * qx; t3 = qw * qy; t4 = qw * qz; t5 = -qx * qx; t6 = qx * qy; t7 = qx * qz; t8 = -qy * qy; t9 = qy * qz; t10 = -qz * qz;
    d1 = t8 + t10; d2 = t6 - t4; d3 = t3 + t7; d4 = t4 + t6; d5 = t5 + t10; d6 = t9 - t2; d7 = t7 - t3; d8 = t2 + t9; d9 = t5 + t8;
    idx[nv++] = 3;
    if (geotype != SCO_SCN && geotype != SCO_SCA) {
        m = is_two_patch(geotype) ? 2 : 1;
        for (k = 0; k < m; k++) { idx[nv++] = 6 + 3 * k; idx[nv++] = 12 + 6 * k; idx[nv++] = 15 + 6 * k; }
    }
    if (is_chiral(geotype)) {
        m = (geotype == SCO_TCHPSC || geotype == SCO_TCHCPSC) ? 2 : 1;
        for (k = 0; k < m; k++) idx[nv++] = 24 + 3 * k;
    }
    for (k = 0; k < nv; k++) {
        double* p = st + idx[k];
        double nx = 2.0 * (d1 * p[0] + d2 * p[1] + d3 * p[2]) + p[0];
        double ny = 2.0 * (d4 * p[0] + d5 * p[1] + d6 * p[2]) + p[1];
        double nz = 2.0 * (d7 * p[0] + d8 * p[1] + d9 * p[2]) + p[2];
        p[0] = nx; p[1] = ny; p[2] = nz;
    }
}

/* Conf::linemin, structures/Conf.h:461-465 */
static inline double linemin(double criterion, double halfl) {
    if (criterion >= halfl) return halfl;
    else if (criterion >= -halfl) return criterion;
    else return -halfl;
}

/* Conf::overlap, structures/Conf.cpp:104-239.
 * variant 0 = as written: spheres keyed on the type NUMBER >= 30, rod-rod half length = half_len[1]/2,
 * threshold sigma/2. variant 1 = documented intent: geotype-keyed, half_len[0], threshold sigma. */
int sco_overlap_pair(const sco_system* s, const double* s1, int type1, const double* s2, int type2, int variant) {
    const sco_iaparam* ia = &s->ia[type1 * s->ntypes + type2];
    vec3 r_cm = image(s->box, ld(s1), ld(s2));
    vec3 dir1 = ld(s1 + 3), dir2 = ld(s2 + 3), distvec;
    double b, c, d, e, f, boundary, det, halfl, s0, t0, ss, tt, dist;
    int g0 = (int)ia->geotype[0], g1 = (int)ia->geotype[1];
    int both_spheres = variant == 0 ? ((type1 >= SCO_SP) && (type2 >= SCO_SP)) : (g0 >= SCO_SP && g1 >= SCO_SP);
    if (both_spheres) {
        dist = sqrt(dot(r_cm, r_cm));
    } else if ((g0 < SCO_SP) && (g1 < SCO_SP)) {
        b = -dot(dir1, dir2);
        d = dot(dir1, r_cm);
        e = -dot(dir2, r_cm);
        f = dot(r_cm, r_cm);
        det = 1.0 - b * b;
        halfl = ia->half_len[1];
        if (variant == 0) halfl /= 2;
        boundary = det * halfl;
        s0 = b * e - d;
        t0 = b * d - e;
        if (s0 >= boundary) {
            if (t0 >= boundary) {
                if (d + halfl + halfl * b < 0.0) { ss = halfl; tt = linemin(-ss * b - e, halfl); }
                else { tt = halfl; ss = linemin(-tt * b - d, halfl); }
            } else if (t0 >= -boundary) {
                ss = halfl; tt = linemin(-ss * b - e, halfl);
            } else {
                if (d + halfl - halfl * b < 0.0) { ss = halfl; tt = linemin(-ss * b - e, halfl); }
                else { tt = -halfl; ss = linemin(-tt * b - d, halfl); }
            }
        } else if (s0 >= -boundary) {
            if (t0 >= boundary) { tt = halfl; ss = linemin(-tt * b - d, halfl); }
            else if (t0 >= -boundary) { ss = s0 / det; tt = t0 / det; }
            else { tt = -halfl; ss = linemin(-tt * b - d, halfl); }
        } else {
            if (t0 >= boundary) {
                if (d - halfl + halfl * b > 0.0) { ss = -halfl; tt = linemin(-ss * b - e, halfl); }
                else { tt = halfl; ss = linemin(-tt * b - d, halfl); }
            } else if (t0 >= -boundary) {
                ss = -halfl; tt = linemin(-ss * b - e, halfl);
            } else {
                if (d - halfl - halfl * b > 0.0) { ss = -halfl; tt = linemin(-ss * b - e, halfl); }
                else { tt = -halfl; ss = linemin(-tt * b - d, halfl); }
            }
        }
        dist = sqrt(f + ss * ss + tt * tt + 2.0 * (ss * d + tt * e + ss * tt * b));
    } else if (g0 < SCO_SP) {
        halfl = ia->half_len[0];
        c = dot(dir1, r_cm);
        if (c >= halfl) d = halfl; else { if (c > -halfl) d = c; else d = -halfl; }
        distvec = V(-r_cm.x + dir1.x * d, -r_cm.y + dir1.y * d, -r_cm.z + dir1.z * d);
        dist = sqrt(dot(distvec, distvec));
    } else {
        halfl = ia->half_len[1];
        c = dot(dir2, r_cm);
        if (c >= halfl) d = halfl; else { if (c > -halfl) d = c; else d = -halfl; }
        distvec = V(r_cm.x - dir2.x * d, r_cm.y - dir2.y * d, r_cm.z - dir2.z * d);
        dist = sqrt(dot(distvec, distvec));
    }
    if (dist < ia->sigma * (variant == 0 ? 0.5 : 1.0)) return 1;
    return 0;
}

/* Conf::overlapAll, structures/Conf.cpp:244-253 */
int sco_overlap_one(const sco_system* s, int target, const double* trial_state, int variant) {
    const double* st = trial_state ? trial_state : s->state + (size_t)target * SCO_STATE;
    int i;
    for (i = 0; i < s->n; i++)
        if (i != target && sco_overlap_pair(s, st, s->type[target], s->state + (size_t)i * SCO_STATE, s->type[i], variant)) return 1;
    return 0;
}

/* Conf::checkall, structures/Conf.cpp:256-267 */
int sco_overlap_all(const sco_system* s, int variant) {
    int i, j;
    for (i = 0; i + 1 < s->n; i++)
        for (j = i + 1; j < s->n; j++)
            if (sco_overlap_pair(s, s->state + (size_t)i * SCO_STATE, s->type[i], s->state + (size_t)j * SCO_STATE, s->type[j], variant)) return 1;
    return 0;
}

/* ---- cell list, definition C1 (binning convention of Mesh::addPart, mc/mesh.cpp:49-54,
 * and INBOX, structures/macros.h:119) ---- */
void sco_cell_dims(const sco_system* s, int ncell[3]) {
    int d;
    for (d = 0; d < 3; d++) {
        int nc = (int)floor(s->box[d] / s->maxcut);
        if (nc < 3) nc = 1; /* fewer than 3 cells: neighbours would alias through the periodic image */
        ncell[d] = nc;
    }
}

static inline int cell_coord(double u, int nc) {
    double ip, f;
    int c;
    f = (u > 0) ? modf(u, &ip) : modf(u, &ip) + 1;
    c = (int)(f * nc);
    if (c == nc) c = 0;
    return c;
}

void sco_cell_assign(const sco_system* s, int* cell_of, int ncell[3]) {
    int i;
    sco_cell_dims(s, ncell);
    for (i = 0; i < s->n; i++) {
        const double* p = s->state + (size_t)i * SCO_STATE;
        int cx = cell_coord(p[0], ncell[0]), cy = cell_coord(p[1], ncell[1]), cz = cell_coord(p[2], ncell[2]);
        cell_of[i] = (cz * ncell[1] + cy) * ncell[0] + cx;
    }
}

void sco_cell_sort(const sco_system* s, const int* cell_of, int ncells, int* order, int* cell_start) {
    int* cnt = (int*)calloc((size_t)ncells + 1, sizeof(int));
    int i, c;
    for (i = 0; i < s->n; i++) cnt[cell_of[i] + 1]++;
    for (c = 0; c < ncells; c++) cnt[c + 1] += cnt[c];
    if (cell_start) memcpy(cell_start, cnt, ((size_t)ncells + 1) * sizeof(int));
    for (i = 0; i < s->n; i++) order[cnt[cell_of[i]]++] = i; /* stable: ascending original index in a cell */
    free(cnt);
}

static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

double sco_one_to_all_cells(const sco_system* s, int target, const double* trial_state,
                            const int* cell_of, const int ncell[3], const int* order, const int* cell_start,
                            long* n_candidates, long* n_gated) {
    const double* st = trial_state ? trial_state : s->state + (size_t)target * SCO_STATE;
    sco_conlist cl;
    int cx, cy, cz, dx, dy, dz, k, nl = 0, cap = 1024;
    int* list = (int*)malloc(sizeof(int) * cap);
    double energy = 0.0;
    long gated = 0;
    (void)cell_of;
    sco_get_conlist(s, target, &cl);
    cx = cell_coord(st[0], ncell[0]); cy = cell_coord(st[1], ncell[1]); cz = cell_coord(st[2], ncell[2]);
    for (dz = -1; dz <= 1; dz++) {
        if (ncell[2] == 1 && dz != 0) continue;
        for (dy = -1; dy <= 1; dy++) {
            if (ncell[1] == 1 && dy != 0) continue;
            for (dx = -1; dx <= 1; dx++) {
                int c;
                if (ncell[0] == 1 && dx != 0) continue;
                c = (((cz + dz + ncell[2]) % ncell[2]) * ncell[1] + ((cy + dy + ncell[1]) % ncell[1])) * ncell[0] + ((cx + dx + ncell[0]) % ncell[0]);
                for (k = cell_start[c]; k < cell_start[c + 1]; k++) {
                    if (order[k] == target) continue;
                    if (nl == cap) { cap *= 2; list = (int*)realloc(list, sizeof(int) * cap); }
                    list[nl++] = order[k];
                }
            }
        }
    }
    /* A non-empty conlist disables the cutoff gate for EVERY pair of this particle (paire.h:1214), but
     * beyond sqmaxcut every non-bonded term is zero by construction of maxcut (topo.cpp:140-153), so
     * only the (up to 4) bonded partners can contribute from outside the neighbourhood: add them. */
    if (!cl.is_empty) {
        int q, m, have;
        for (q = 0; q < 4; q++) {
            if (cl.con[q] < 0) continue;
            have = 0;
            for (m = 0; m < nl; m++) if (list[m] == cl.con[q]) { have = 1; break; }
            if (!have) {
                if (nl == cap) { cap *= 2; list = (int*)realloc(list, sizeof(int) * cap); }
                list[nl++] = cl.con[q];
            }
        }
    }
    qsort(list, (size_t)nl, sizeof(int), cmp_int);
    for (k = 0; k < nl; k++) {
        int j = list[k];
        vec3 r;
        /* work counter of the cell path: pairs that reach a functor there = inside sqmaxcut, or bonded. The separation is
         * computed again inside sco_pair_energy; the op-counting build must see it once (PairE::operator() does it once). */
        SCO_COUNT_PAUSE();
        r = image(s->box, ld(st), ld(s->state + (size_t)j * SCO_STATE));
        if (dot(r, r) <= s->sqmaxcut || j == cl.con[0] || j == cl.con[1] || j == cl.con[2] || j == cl.con[3]) gated++;
        SCO_COUNT_RESUME();
        energy += sco_pair_energy(s, st, s->type[target], s->moltype[target], target,
                                  s->state + (size_t)j * SCO_STATE, s->type[j], j, &cl);
    }
    if (n_candidates) *n_candidates = nl;
    if (n_gated) *n_gated = gated;
    free(list);
    return energy;
}
