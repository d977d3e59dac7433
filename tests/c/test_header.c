/* The drop-in boundary is a C ABI: include/herbgpu.h must compile as plain C11 and link against libherbgpu.so
 * without any C++ or CUDA header on the consumer's side.  Exercises the argument-checking paths that need no GPU. */
#include <stdio.h>
#include <string.h>

#include "../../include/herbgpu.h"

#define EXPECT(c)                                                        \
    do {                                                                 \
        if (!(c)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } \
    } while (0)

_Static_assert(sizeof(hb_chunk_pt) == 12, "ChunkPt");
_Static_assert(sizeof(hb_instance3d) == 100, "Instance3d");
_Static_assert(sizeof(hb_palette_cube) == 8, "PaletteCube");
_Static_assert(sizeof(hb_chunk_cube) == 12, "ChunkCube");
_Static_assert(sizeof(hb_region) == 24, "hb_region");

int main(void) {
    int failures = 0;
    EXPECT(strstr(hb_version(), "sm_100a") != NULL);
    EXPECT(hb_create(0, NULL) == HB_ERR_INVALID);
    EXPECT(strlen(hb_last_error()) > 0);
    EXPECT(hb_set_world(NULL, 0, 0.001f, 6, 2.0f, 0.6f) == HB_ERR_INVALID);
    EXPECT(hb_set_noise_kind(NULL, HB_NOISE_RIDGE) == HB_ERR_INVALID);
    EXPECT(hb_generate(NULL, NULL, 0, NULL, NULL, NULL) == HB_ERR_INVALID);
    EXPECT(hb_world_create(NULL, 16, NULL) == HB_ERR_INVALID);
    EXPECT(hb_world_len(NULL) == 0);
    EXPECT(hb_world_load(NULL, NULL, 0) == HB_ERR_INVALID);
    EXPECT(hb_world_sync(NULL, NULL, NULL) == HB_ERR_INVALID);
    EXPECT(hb_rle_decode(NULL, NULL, 0, NULL, NULL, 0, NULL) == HB_ERR_INVALID);
    {
        /* pure host helper: shell order of create_chunk_shell, centre = 13 */
        hb_chunk_pt pts[2] = {{0, 0, 0}, {1, 0, 0}};
        int32_t nbr[54];
        EXPECT(hb_build_neighbor_index(pts, 2, nbr) == HB_OK);
        EXPECT(nbr[13] == 0 && nbr[27 + 13] == 1);
        EXPECT(nbr[22] == 1);      /* (+1, 0, 0) of chunk 0 */
        EXPECT(nbr[27 + 4] == 0);  /* (-1, 0, 0) of chunk 1 */
    }
    hb_destroy(NULL);
    hb_world_destroy(NULL);
    hb_region_plan_destroy(NULL);
    printf(failures ? "FAILED\n" : "OK\n");
    return failures ? 1 : 0;
}
