/* TEST INFRASTRUCTURE -- CPU first-hit ray casting from one origin (row N3's oracle).
 *
 * PARITY UNPINNED against the reference: /root/reference/scene/avatar_gaussian_model.py:227-263 and
 * /root/reference/inference.py:285-316 call open3d's RaycastingScene.cast_rays (Intel Embree, closest hit,
 * two-sided, INVALID_ID when nothing is hit); open3d / Embree are not in this image, so there is nothing to run or
 * to take golden vectors from.  This file restates the published semantics of that call -- for every ray the
 * triangle with the smallest positive hit distance along  ray_d = (x_i - camera) / |x_i - camera|  -- with the
 * Moller-Trumbore intersection test, brute force over all triangles, ties to the lower triangle index.  Known and
 * accepted difference: Embree's watertight edge rules can pick the other triangle when a ray passes exactly through
 * a shared edge.  The CUDA path (gaussian-garments_b200/csrc/visibility.cu, built with -fmad=false) uses the same
 * expression order, so the two sides agree bit for bit.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.                               */
#include <math.h>
#include <stdint.h>

/* barycentric slack: a ray through a shared edge must hit at least one of the two triangles (Embree is watertight;
 * plain Moller-Trumbore can reject both by one ulp).  Same constant in csrc/visibility.cu. */
#define RAY_EDGE_EPS 1e-6f

static int ray_tri(float ox, float oy, float oz, float dx, float dy, float dz, const float* v, int64_t i0, int64_t i1,
                   int64_t i2, float* t_out) {
    const float ax = v[3 * i0], ay = v[3 * i0 + 1], az = v[3 * i0 + 2];
    const float e1x = v[3 * i1] - ax, e1y = v[3 * i1 + 1] - ay, e1z = v[3 * i1 + 2] - az;
    const float e2x = v[3 * i2] - ax, e2y = v[3 * i2 + 1] - ay, e2z = v[3 * i2 + 2] - az;
    const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
    const float det = e1x * px + e1y * py + e1z * pz;
    if (det == 0.f) return 0;
    const float inv = 1.0f / det;
    const float tx = ox - ax, ty = oy - ay, tz = oz - az;
    const float bu = (tx * px + ty * py + tz * pz) * inv;
    if (bu < -RAY_EDGE_EPS || bu > 1.f + RAY_EDGE_EPS) return 0;
    const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    const float bv = (dx * qx + dy * qy + dz * qz) * inv;
    if (bv < -RAY_EDGE_EPS || bu + bv > 1.f + RAY_EDGE_EPS) return 0;
    const float t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    if (!(t > 0.f)) return 0;
    *t_out = t;
    return 1;
}

/* prim[i] = index of the first triangle hit by the ray origin -> targets[i] (-1: none); t_hit[i] = its distance (inf). */
int ggo_cast_rays_from_point(int32_t V, int32_t F, int32_t N, const float* verts, const int32_t* faces,
                             const float* targets, const float* origin, int32_t* prim, float* t_hit) {
    (void)V;
    const float ox = origin[0], oy = origin[1], oz = origin[2];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        float dx = targets[3 * (int64_t)i] - ox, dy = targets[3 * (int64_t)i + 1] - oy, dz = targets[3 * (int64_t)i + 2] - oz;
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
        dx /= nrm; dy /= nrm; dz /= nrm;
        int best = -1;
        float best_t = 3.0e38f;
        for (int f = 0; f < F; f++) {
            float t;
            if (ray_tri(ox, oy, oz, dx, dy, dz, verts, faces[3 * (int64_t)f], faces[3 * (int64_t)f + 1], faces[3 * (int64_t)f + 2], &t) &&
                (t < best_t || (t == best_t && f < best))) {
                best_t = t;
                best = f;
            }
        }
        prim[i] = best;
        if (t_hit) t_hit[i] = best >= 0 ? best_t : INFINITY;
    }
    return 0;
}
