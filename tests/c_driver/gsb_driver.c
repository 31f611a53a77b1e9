/* gsb_driver.c — a Houdini stand-in in plain C: drives libgsplat_b200.so through its C ABI (include/gsplat_b200.h) with
 * the call sequence of the reference's hooks, no Python and no ctypes in between:
 *
 *   GR_PrimGsplat::update  -> registerUpdate                      (src/GR_GSplat.C:423-436)      gsb_register_update
 *   GR_PrimGsplat::render  -> includeInRenderPass, setSphericalHarmonicsOrder (GR_GSplat.C:485-492)
 *   MyCustomSceneRenderHook::render -> generateRenderGeometry, render, postRender (DM_GSplatHook.C:30-39)
 *   GR_PrimGsplat::~GR_PrimGsplat -> flushEntriesForMatchingDetail (GR_GSplat.C:63-70)
 *
 * usage: gsb_driver <scene.bin> <frame_out.bin> [frames]
 * scene.bin (little endian, written by tests/test_c_driver.py): int64 n; int32 has_sh; int32 sh_order; float origin[3];
 *   int32 pad; gsb_frame (as laid out by the header); float pos[n][3]; uint16 cd[n][3]; float alpha[n]; uint16 scale[n][3];
 *   uint16 orient[n][4]; if has_sh: uint16 shx[n][16], shy[n][16], shz[n][16].
 * frame_out.bin: width*height*4 floats (row 0 = bottom scanline) followed by the gsb_stats struct of the last frame.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "gsplat_b200.h"

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != GSB_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, gsb_last_error()); return 10 + rc_; } \
    } while (0)

static void* read_block(FILE* f, size_t bytes)
{
    void* p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read (%zu bytes)\n", bytes); exit(3); }
    return p;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <scene.bin> <frame_out.bin> [frames]\n", argv[0]); return 2; }
    const int frames = argc > 3 ? atoi(argv[3]) : 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int64_t n; int32_t has_sh, sh_order, pad; float origin[3]; gsb_frame frame;
    if (fread(&n, 8, 1, f) != 1 || fread(&has_sh, 4, 1, f) != 1 || fread(&sh_order, 4, 1, f) != 1 ||
        fread(origin, 4, 3, f) != 3 || fread(&pad, 4, 1, f) != 1 || fread(&frame, sizeof frame, 1, f) != 1) {
        fprintf(stderr, "bad scene header\n"); return 3;
    }
    const size_t N = (size_t)n;
    float* pos = read_block(f, N * 12); uint16_t* cd = read_block(f, N * 6); float* alpha = read_block(f, N * 4);
    uint16_t* scale = read_block(f, N * 6); uint16_t* orient = read_block(f, N * 8);
    uint16_t *shx = NULL, *shy = NULL, *shz = NULL;
    if (has_sh) { shx = read_block(f, N * 32); shy = read_block(f, N * 32); shz = read_block(f, N * 32); }
    fclose(f);

    if (gsb_abi_version() != GSB_ABI_VERSION) { fprintf(stderr, "ABI mismatch: library %d, header %d\n", gsb_abi_version(), GSB_ABI_VERSION); return 4; }
    gsb_context* ctx = NULL;
    CHECK(gsb_create(0, &ctx));                                       /* GSplatRenderer::getInstance() */

    gsb_prim_key key = { 0x7f00c0de0000ull, 0, { 1, 0, 0, 0 } };
    char id[GSB_ID_MAX];
    CHECK(gsb_register_update(ctx, &key, n, origin, pos, cd, alpha, scale, orient, shx, shy, shz, id));

    const size_t px = (size_t)frame.width * (size_t)frame.height;
    float* rgba = calloc(px * 4, sizeof(float));
    gsb_target target; memset(&target, 0, sizeof target);
    target.host_rgba = rgba;
    gsb_stats st; memset(&st, 0, sizeof st);
    for (int k = 0; k < frames; ++k) {                                /* one viewport redraw per iteration */
        CHECK(gsb_set_rendering_enabled(ctx, 1));                     /* GR_GSplat.C:472 */
        CHECK(gsb_include_in_render_pass(ctx, id));                   /* GR_GSplat.C:485 */
        CHECK(gsb_set_spherical_harmonics_order(ctx, sh_order));      /* GR_GSplat.C:492 */
        CHECK(gsb_generate_render_geometry(ctx));                     /* DM_GSplatHook.C:32 */
        CHECK(gsb_render(ctx, &frame, &target));                      /* DM_GSplatHook.C:34 */
        CHECK(gsb_post_render(ctx));                                  /* DM_GSplatHook.C:36 */
        CHECK(gsb_get_stats(ctx, &st));
    }
    if (!st.rendered) { fprintf(stderr, "nothing rendered\n"); return 5; }

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    fwrite(rgba, sizeof(float), px * 4, o);
    fwrite(&st, sizeof st, 1, o);
    fclose(o);
    printf("gsb_driver: id=%s N=%lld V=%lld L=%lld D=%lld D_c=%lld launches=%d chunks=%d\n", id, (long long)st.n_submitted,
           (long long)st.n_visible, (long long)st.n_live, (long long)st.n_instances, (long long)st.n_consumed, st.launches, st.depth_chunks);

    CHECK(gsb_flush_entries_for_matching_detail(ctx, id));            /* GR_GSplat.C:67 */
    if (gsb_registry_size(ctx) != 0) { fprintf(stderr, "registry not empty after flush\n"); return 6; }
    CHECK(gsb_destroy(ctx));
    free(pos); free(cd); free(alpha); free(scale); free(orient); free(shx); free(shy); free(shz); free(rgba);
    return 0;
}
