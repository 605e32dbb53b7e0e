// Batched LoS / blockage test of links against the walls of a city.
//
// Replaces the per-link MATLAB loops of networkTopology.blockages.openStreetMapCity.checkLoS
// (+networkTopology/+blockages/openStreetMapCity.m:67-95; call sites networkSimulation.m:138,154: one call per UE and
// per target, each looping over every building and wall) by one launch over all links:
//   building.checkBlockage   (+networkTopology/+blockages/building.m:113-137): a building blocks when any wall does;
//   wallBlockage.checkBlockage (wallBlockage.m:121-127): the user is projected onto the wall plane ALONG THE LINK
//       projUe = ue + (ue-ant) * (normDist - n'ue) / (n'(ue-ant))      (infinite line: reference behaviour kept)
//   wallBlockage.getWindingNumber (wallBlockage.m:178-222): unit vectors point->corners, signed angles between
//       consecutive ones about the wall normal, |sum| > 0.1 -> inside the polygon -> blocked; a point within 1e-10 of a
//       corner counts as inside.
// float64 throughout (the decision is a threshold on a sum of atan2 values; the reference computes in double).
// One CTA per link, threads stride over the walls, block-wide OR.  The plane of every wall (normal, distance) is computed
// once on the host when the city is created (wallBlockage.m:62-71).
#include "los.cuh"
#include "ctx.cuh"
#include <cmath>

namespace isac {

__global__ void __launch_bounds__(128)
check_los_kernel(const double* __restrict__ corners, const int* __restrict__ wallOff, const double* __restrict__ plane, int nWalls,
                 const double* __restrict__ ue, const double* __restrict__ ant, int antStride, int* __restrict__ los) {
    const int link = blockIdx.x;
    const double ux = ue[3 * link], uy = ue[3 * link + 1], uz = ue[3 * link + 2];
    const double* a = ant + (long long)antStride * link;
    const double vx = ux - a[0], vy = uy - a[1], vz = uz - a[2];   // vec = ue - ant        (wallBlockage.m:121)
    int blocked = 0;
    for (int w = threadIdx.x; w < nWalls; w += blockDim.x) {
        const double nx = plane[4 * w], ny = plane[4 * w + 1], nz = plane[4 * w + 2], nd = plane[4 * w + 3];
        const double t = (nd - (nx * ux + ny * uy + nz * uz)) / (nx * vx + ny * vy + nz * vz);
        const double px = ux + vx * t, py = uy + vy * t, pz = uz + vz * t;   // projUe        (:123)
        const int c0 = wallOff[w], c1 = wallOff[w + 1];
        // unit vector to the LAST corner first: shiftvec = circshift(vec,1,3) pairs corner k with corner k-1 (:209)
        double qx = corners[3 * (c1 - 1)] - px, qy = corners[3 * (c1 - 1) + 1] - py, qz = corners[3 * (c1 - 1) + 2] - pz;
        double ql = sqrt(qx * qx + qy * qy + qz * qz);
        bool invalid = ql < 1e-10;                                        // :203
        qx /= ql; qy /= ql; qz /= ql;                                     // :207
        double sum = 0.0;
        for (int c = c0; c < c1; ++c) {
            double rx = corners[3 * c] - px, ry = corners[3 * c + 1] - py, rz = corners[3 * c + 2] - pz;
            const double rl = sqrt(rx * rx + ry * ry + rz * rz);
            invalid |= rl < 1e-10;
            rx /= rl; ry /= rl; rz /= rl;
            const double dotv = qx * rx + qy * ry + qz * rz;              // dot(shiftvec,vec)       (:210)
            const double cx = qy * rz - qz * ry, cy = qz * rx - qx * rz, cz = qx * ry - qy * rx;   // cross (:211)
            sum += atan2(nx * cx + ny * cy + nz * cz, dotv);              // diffAngle, summed in corner order (:215,:220)
            qx = rx; qy = ry; qz = rz;
        }
        double wn = fabs(sum);
        if (invalid) wn = 1.0;                                            // :222
        if (wn > 0.1) blocked = 1;                                        // :127 (NaN compares false -> not blocked)
    }
    blocked = __syncthreads_or(blocked);
    if (threadIdx.x == 0) los[link] = blocked ? 0 : 1;                    // losDecision = ~blockageDecision (openStreetMapCity.m:93)
}

// plane of a wall (wallBlockage.m:62-71).  The reference takes an orthonormal basis of span{c1 - ck} (orth) and crosses
// its two vectors; for a planar polygon that is +- the unit plane normal.  Here: Newell's method (area-weighted normal of
// the polygon, robust to collinear leading corners), normalised; the sign is irrelevant (abs of the winding sum).
static bool wall_plane(const double* c, int n, double out[4]) {
    double nx = 0, ny = 0, nz = 0;
    for (int i = 0; i < n; ++i) {
        const double* p = c + 3 * i;
        const double* q = c + 3 * ((i + 1) % n);
        nx += (p[1] - q[1]) * (p[2] + q[2]);
        ny += (p[2] - q[2]) * (p[0] + q[0]);
        nz += (p[0] - q[0]) * (p[1] + q[1]);
    }
    const double l = std::sqrt(nx * nx + ny * ny + nz * nz);
    if (!(l > 0)) return false;
    out[0] = nx / l; out[1] = ny / l; out[2] = nz / l;
    out[3] = out[0] * c[0] + out[1] * c[1] + out[2] * c[2];               // normDist = normVec' * cornerList(:,1)
    return true;
}

int city_plan_create(Ctx* ctx, int nWalls, const int* wallOff, const double* corners, CityPlan** out) {
    if (nWalls < 1 || !wallOff || !corners || wallOff[0] != 0) {
        set_error(ctx, "city: invalid wall list");
        return kErrInvalidArg;
    }
    std::vector<double> plane((size_t)4 * nWalls);
    for (int w = 0; w < nWalls; ++w) {
        const int n = wallOff[w + 1] - wallOff[w];
        if (n < 3) {  // wallBlockage.m:36-38
            set_error(ctx, "city: use at least three points to specify a wall");
            return kErrInvalidArg;
        }
        if (!wall_plane(corners + 3 * (size_t)wallOff[w], n, &plane[(size_t)4 * w])) {
            set_error(ctx, "city: degenerate wall (corners are collinear)");
            return kErrInvalidArg;
        }
    }
    CityPlan* p = new CityPlan();
    p->ctx = ctx;
    p->nWalls = nWalls;
    p->nCorners = wallOff[nWalls];
    cudaError_t e = cudaMalloc((void**)&p->d_corners, sizeof(double) * 3 * (size_t)p->nCorners);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_wallOff, sizeof(int) * ((size_t)nWalls + 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_plane, sizeof(double) * 4 * (size_t)nWalls);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_corners, corners, sizeof(double) * 3 * (size_t)p->nCorners, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_wallOff, wallOff, sizeof(int) * ((size_t)nWalls + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_plane, plane.data(), sizeof(double) * 4 * (size_t)nWalls, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error(ctx, std::string("city: ") + cudaGetErrorString(e));
        city_plan_destroy(p);
        return kErrCuda;
    }
    *out = p;
    return kOk;
}

void city_plan_destroy(CityPlan* p) {
    if (!p) return;
    cudaFree(p->d_corners);
    cudaFree(p->d_wallOff);
    cudaFree(p->d_plane);
    delete p;
}

int city_check_los(CityPlan* p, int n, const double* ue, const double* ant, int antStride, int* los, cudaStream_t st) {
    Ctx* ctx = p->ctx;
    if (n < 1 || !ue || !ant || !los || (antStride != 0 && antStride != 3)) {
        set_error(ctx, "checkLoS: invalid argument");
        return kErrInvalidArg;
    }
    check_los_kernel<<<n, 128, 0, st>>>(p->d_corners, p->d_wallOff, p->d_plane, p->nWalls, ue, ant, antStride, los);
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

}  // namespace isac
