/* ORACLE — test infrastructure only.  Never imported by the product path
 * (sa-ssd_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * CPU restatement of the reference hard voxelizer:
 *   mmdet/ops/points_op/points_ops.py:4-50   (_points_to_voxel_reverse_kernel)
 *   mmdet/ops/points_op/points_ops.py:104-164 (points_to_voxel: buffers, dense
 *                                              coor_to_voxelidx table, slicing)
 * Pinned against the reference's own numba kernel executed in the build
 * container (tests/golden/make_golden.py -> tests/golden/voxelize_*.npz).
 *
 * Semantics reproduced exactly: fp32 (p - lo) / vs then floor; reject a point
 * when any coordinate leaves [0, grid); voxel ids by first touch; first
 * max_points points kept per voxel; processing of ALL remaining points stops
 * at the first point that would open voxel number max_voxels (points_ops.py:41-42).
 * Build with -ffp-contract=off and without -ffast-math.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* points [N, ndim] f32; voxel_size [3] (x,y,z); range [6] (xyzxyz min max).
 * outputs sized for max_voxels: voxels [max_voxels, max_points, ndim] (must be
 * zero-filled by the caller or are zeroed here), coors [max_voxels, 3] (z,y,x),
 * num_points [max_voxels].  Returns the number of voxels, or -1 on allocation
 * failure. */
int oracle_points_to_voxel(const float *points, int n, int ndim,
                           const float *voxel_size, const float *range,
                           int max_points, int max_voxels,
                           float *voxels, int32_t *coors, int32_t *num_points)
{
    int32_t grid[3];
    for (int j = 0; j < 3; ++j) {
        /* points_ops.py:21-24: (range[3:] - range[:3]) / voxel_size, np.round */
        float g = (range[3 + j] - range[j]) / voxel_size[j];
        grid[j] = (int32_t)nearbyintf(g);
    }
    /* dense (z, y, x) table, -1 = empty: points_ops.py:145 */
    size_t cells = (size_t)grid[0] * (size_t)grid[1] * (size_t)grid[2];
    int32_t *table = (int32_t *)malloc(cells * sizeof(int32_t));
    if (!table) return -1;
    memset(table, 0xff, cells * sizeof(int32_t));
    memset(voxels, 0, (size_t)max_voxels * max_points * ndim * sizeof(float));
    memset(num_points, 0, (size_t)max_voxels * sizeof(int32_t));
    memset(coors, 0, (size_t)max_voxels * 3 * sizeof(int32_t));

    int voxel_num = 0;
    for (int i = 0; i < n; ++i) {
        int32_t coor[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            float c = floorf((points[(size_t)i * ndim + j] - range[j]) / voxel_size[j]);
            if (c < 0 || c >= (float)grid[j]) { failed = 1; break; }
            coor[2 - j] = (int32_t)c;
        }
        if (failed) continue;
        size_t cell = ((size_t)coor[0] * grid[1] + coor[1]) * grid[0] + coor[2];
        int32_t vid = table[cell];
        if (vid == -1) {
            vid = voxel_num;
            if (voxel_num >= max_voxels) break;
            voxel_num += 1;
            table[cell] = vid;
            coors[vid * 3 + 0] = coor[0];
            coors[vid * 3 + 1] = coor[1];
            coors[vid * 3 + 2] = coor[2];
        }
        int32_t num = num_points[vid];
        if (num < max_points) {
            memcpy(voxels + ((size_t)vid * max_points + num) * ndim,
                   points + (size_t)i * ndim, ndim * sizeof(float));
            num_points[vid] = num + 1;
        }
    }
    free(table);
    return voxel_num;
}
