// LoS / blockage geometry of the city layout: declarations shared with capi.cu.
#pragma once
#include "isac_common.cuh"
#include <vector>

namespace isac {

struct CityPlan {
    Ctx* ctx = nullptr;
    int nWalls = 0, nCorners = 0;
    double* d_corners = nullptr;  // [3 x nCorners] column-major (x;y;z per corner), walls concatenated
    int* d_wallOff = nullptr;     // [nWalls + 1] corner offsets
    double* d_plane = nullptr;    // [4 x nWalls]: unit normal (3) and plane distance normVec'*c1
};

// walls: polygons in 3-D (wallBlockage.cornerList), concatenated; wallOff[nWalls+1].
int city_plan_create(Ctx* ctx, int nWalls, const int* wallOff, const double* corners, CityPlan** out);
void city_plan_destroy(CityPlan* p);
// los[i] = 1 when no wall blocks link i (ue_i -> ant_i; antStride = 0 broadcasts one antenna position).
// ue / ant: device double [3 x n]; los: device int32 [n].
int city_check_los(CityPlan* p, int n, const double* ue, const double* ant, int antStride, int* los, cudaStream_t st);

}  // namespace isac
