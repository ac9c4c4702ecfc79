// External wall potential ([EXTER] section of top.init): ExternalEnergyCalculator::extere2 and its helpers
// (scOOP/mc/externalenergycalculator.cpp:5-500, externalenergycalculator.h:21-115), one particle against a structureless wall in the
// plane z = 0 of the periodic box: WCA repulsion of the closest point, attraction of the patch through the projection of the
// interacting piece of the rod on the wall. The reference keeps the intermediate values in members of the calculator object
// (dist, positive, orientin, orient, rcmz, interendz, project); here they are the fields of WallState, same names, and every function
// follows the operation order of its original so that the strict build agrees with it to the last bit away from libm calls.
// project is (0, 0, +-1) throughout extere2, so projectinZ() is the orthogonal projection on the wall; it is kept as written.
#pragma once
#include "pair_energy.cuh"

namespace scg {

struct WallParam {          // topo.exter.interactions[type] (structures/topo.cpp:120-130): the type's own Ia_param with the wall's mixing rules
    double sigma, epsilon, rcutwca, rcut, pdis, pswitch;
    double len0, half_len0;
    int geotype;
    int pad;
};

#define SCG_ZEROTOL2 1.0e-8        // ZEROTOL2 (structures/macros.h)
#define SCG_AVER(a, b) (((a) + (b)) * 0.5)

struct WallState {
    double dist, orient, rcmz, interendz;
    bool positive, orientin;
    v3 project;
    v3 dir;                 // part1->dir (replaced by chdir[k] for chiral types, externalenergycalculator.cpp:66-69, 82-85)
};

__device__ inline void wall_project_in_z(const v3& vec1, const v3& projectdir, v3& projection) {
    projection.x = vec1.x - vec1.z * projectdir.x / projectdir.z;
    projection.y = vec1.y - vec1.z * projectdir.y / projectdir.z;
    projection.z = 0;
}

// ExternalEnergyCalculator::exter2ClosestDist (:107-145)
__device__ inline void wall_closest_dist(WallState& w, const WallParam& param) {
    if (w.rcmz < 0) { w.dist = -(w.rcmz); w.positive = false; w.interendz = -1.0; w.project.z = 1.0; }
    else { w.dist = (w.rcmz); w.positive = true; w.interendz = 1.0; w.project.z = -1.0; }
    if (param.geotype < SCGPU_SPN) {        // psc closest is always the end closer to the wall
        if (w.dir.z > 0) {
            if (w.positive) { w.orientin = false; w.orient = -1.0; w.dist = w.rcmz - w.dir.z * param.half_len0; }
            else { w.orientin = true; w.orient = 1.0; w.dist = -(w.rcmz + w.dir.z * param.half_len0); }
        } else {
            if (w.positive) { w.orientin = true; w.orient = 1.0; w.dist = w.rcmz + w.dir.z * param.half_len0; }
            else { w.orientin = false; w.orient = -1.0; w.dist = -(w.rcmz - w.dir.z * param.half_len0); }
        }
    }
}

// ExternalEnergyCalculator::pscWall (:263-458)
__device__ inline int wall_psc(const WallState& w, v3& pbeg, v3& pend, const v3& projectdir, const v3& partdir, double cutdist,
                                        const v3& partbeg, const v3& partend) {
    v3 vec1;
    double k, x1, x2, y1, y2, a, b, c, e, d;
    if (((w.positive) && (projectdir.z > 0)) || ((!(w.positive)) && (projectdir.z < 0))) return 0;
    if (fabs(partbeg.z) > cutdist) return 0;
    x2 = 0.0;
    y2 = 0.0;
    if (fabs(partdir.z) > SCG_ZEROTOL2) { wall_project_in_z(partbeg, partdir, pbeg); a = 0; }
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
        if (fabs(projectdir.z) > SCG_ZEROTOL2) wall_project_in_z(partbeg, projectdir, pend);
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
    // end point
    a = -cutdist * projectdir.z;      // z coordinate of the point where the projection is at the cut distance
    if (((partend.z < a) && (w.positive)) || ((a < partend.z) && (!(w.positive)))) {
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
        if (((partbeg.z < a) && (w.positive)) || ((a < partbeg.z) && (!(w.positive)))) {
            // the end is at the cutoff, going through the cylindrical part
            b = (a - partbeg.z) / partdir.z;
            vec1.x = partbeg.x + b * partdir.x;
            vec1.y = partbeg.y + b * partdir.y;
            vec1.z = a;
            wall_project_in_z(vec1, projectdir, pend);
        } else {
            // the projected end is within the same sphere as the beginning: no contribution from the cylinder
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

// ExternalEnergyCalculator::cpscWall (:460-500)
__device__ inline int wall_cpsc(const WallState& w, v3& pbeg, v3& pend, const v3& projectdir, const v3& partdir, double halfl,
                                         double cutdist, const v3& partbeg, const v3& partend) {
    v3 vec1;
    double a;
    if (((w.positive) && (projectdir.z >= 0)) || ((!(w.positive)) && (projectdir.z <= 0))) return 0;
    vec1.x = partbeg.x;
    vec1.y = partbeg.y;
    vec1.z = partbeg.z;
    if (-vec1.z / projectdir.z < cutdist) wall_project_in_z(vec1, projectdir, pbeg);
    else return 0;
    if (-partend.z / projectdir.z < cutdist) vec1.z = partend.z;
    else vec1.z = -cutdist * projectdir.z;
    if (partdir.z != 0.0) a = (vec1.z - (w.rcmz)) / partdir.z;
    else { if (w.orientin) a = -halfl; else a = halfl; }
    vec1.x = partdir.x * a;
    vec1.y = partdir.y * a;
    wall_project_in_z(vec1, projectdir, pend);
    return 1;
}

// ExternalEnergyCalculator::exter2Atre (:147-261)
__device__ inline double wall_atre(WallState& w, const WallParam& param, double* ndist, const v3& patchdir, double halfl) {
    v3 pbeg, pend;
    double a, length1, length2, f0, f1;
    v3 cm1, cm2;
    int line;
    v3 partbeg, partend;
    v3 inters;
    double atrenergy = 0.0;
    pbeg = mk(0, 0, 0); pend = mk(0, 0, 0);
    if ((param.geotype < SCGPU_SPN) && (param.geotype > SCGPU_SCA)) {
        a = ((w.orientin ? 1.0 : 0.0) - 0.5) * 2;
        partbeg.x = a * w.dir.x * halfl;
        partbeg.y = a * w.dir.y * halfl;
        partbeg.z = w.rcmz + a * w.dir.z * halfl;
        partend.x = -a * w.dir.x * halfl;
        partend.y = -a * w.dir.y * halfl;
        partend.z = w.rcmz - a * w.dir.z * halfl;
        if ((param.rcut - w.dist) / fabs(w.dir.z) < 2.0 * halfl) w.interendz *= param.rcut;
        else w.interendz = partend.z;
        if (w.positive) cm1.z = SCG_AVER(w.interendz, w.dist);
        else cm1.z = SCG_AVER(w.interendz, -w.dist);
        if (w.dir.z != 0.0) {
            a = (w.interendz - cm1.z) / w.dir.z;
            length1 = -w.orient * 2.0 * a;
            a = a + w.orient * halfl;
        } else {
            a = 0.0;
            length1 = 2.0 * halfl;
        }
        cm1.x = w.dir.x * a;
        cm1.y = w.dir.y * a;
        if ((param.geotype == SCGPU_CPSC) || (param.geotype == SCGPU_CHCPSC)) {
            if (((w.interendz >= w.dist) && (w.positive)) || ((w.interendz <= -w.dist) && (!(w.positive))))
                line = wall_cpsc(w, pbeg, pend, w.project, w.dir, param.half_len0, param.rcut, partbeg, partend);
            else line = 0;
        } else {
            line = wall_psc(w, pbeg, pend, w.project, w.dir, param.rcut, partbeg, partend);
        }
        if (line > 0) {
            cm2.x = SCG_AVER(pbeg.x, pend.x);
            cm2.y = SCG_AVER(pbeg.y, pend.y);
            cm2.z = 0.0;
            length2 = sqrt((pend.x - pbeg.x) * (pend.x - pbeg.x) + (pend.y - pbeg.y) * (pend.y - pbeg.y));
            inters.x = cm2.x - cm1.x;
            inters.y = cm2.y - cm1.y;
            inters.z = cm2.z - cm1.z;
            *ndist = sqrt(inters.x * inters.x + inters.y * inters.y + inters.z * inters.z);
            if (*ndist < param.pdis) atrenergy = -param.epsilon;
            else {
                atrenergy = cos(SCG_PIH * (*ndist - param.pdis) / param.pswitch);
                atrenergy *= -atrenergy * param.epsilon;
            }
            f0 = (length1 + length2) * 0.5;
            f1 = fabs(patchdir.z);
            atrenergy *= f0 * f1;
        } else atrenergy = 0.0;
    } else {
        if (*ndist < param.pdis) atrenergy = -param.epsilon;
        else {
            atrenergy = cos(SCG_PIH * (*ndist - param.pdis) / param.pswitch);
            atrenergy *= -atrenergy * param.epsilon;
        }
        atrenergy *= (param.rcut * param.rcut - (*ndist) * (*ndist)) / (param.sigma * param.sigma);
    }
    return atrenergy;
}

// ExternalEnergyCalculator::extere2 (:5-105). pos_z: box-fractional z; dir, patchdir[2], chdir[2]: the particle's vectors.
__device__ inline double wall_energy(const WallParam& param, double exter_sqmaxcut, double box_z, double pos_z, const v3& dir,
                                              const v3& patchdir0, const v3& patchdir1, const v3& chdir0, const v3& chdir1) {
    double repenergy = 0.0, atrenergy = 0.0;
    double ndist, halfl;
    WallState w;
    if (pos_z < 0) w.rcmz = box_z * (pos_z - (double)((long long)(pos_z - 0.5)));
    else w.rcmz = box_z * (pos_z - (double)((long long)(pos_z + 0.5)));
    w.project = mk(0, 0, 0);
    if (w.rcmz < 0) { w.dist = -w.rcmz; w.positive = false; w.interendz = -1.0; w.project.z = 1.0; }
    else { w.dist = w.rcmz; w.positive = true; w.interendz = 1.0; w.project.z = -1.0; }
    if (w.rcmz * w.rcmz > exter_sqmaxcut) return 0.0;
    halfl = 0.5 * param.len0;
    ndist = w.dist;
    w.orientin = true;
    w.orient = 0.0;
    w.dir = dir;
    wall_closest_dist(w, param);
    if (w.dist > param.rcutwca) repenergy = 0.0;
    else {
        const double q = param.sigma / w.dist;
        const double en6 = (q * q * q) * (q * q * q);       // pow(sigma / dist, 6)
        repenergy = 4 * en6 * (en6 - 1) + 1.0;
    }
    if ((param.geotype == SCGPU_CHCPSC) || (param.geotype == SCGPU_CHPSC)) {
        w.dir = chdir0;
        wall_closest_dist(w, param);
    }
    if ((w.dist > param.rcut) || (param.epsilon == 0.0) || ((patchdir0.z > 0) && (w.positive)) || ((patchdir0.z < 0) && (!w.positive))) atrenergy = 0.0;
    else atrenergy = wall_atre(w, param, &ndist, patchdir0, halfl);
    if ((param.geotype == SCGPU_TCPSC) || (param.geotype == SCGPU_TPSC) || (param.geotype == SCGPU_TCHCPSC) || (param.geotype == SCGPU_TCHPSC)) {
        if ((param.geotype == SCGPU_TCHCPSC) || (param.geotype == SCGPU_TCHPSC)) {
            w.dir = chdir1;
            wall_closest_dist(w, param);
        }
        wall_closest_dist(w, param);
        if ((w.dist > param.rcut) || (param.epsilon == 0.0) || ((patchdir1.z > 0) && (w.positive)) || ((patchdir1.z < 0) && (!(w.positive)))) atrenergy += 0.0;
        else atrenergy += wall_atre(w, param, &ndist, patchdir1, halfl);
    }
    return repenergy + atrenergy;
}

}  // namespace scg
