// TEST SHIM (not a product path): compiles the exact-arithmetic header for the host so that tests/test_hostmath.py can
// compare the very functions the kernels run against the oracle on a machine without a GPU.
#include <string.h>

#include "o2v_exact.cuh"

using namespace o2v;

extern "C" {

int o2vt_clip_voxel(const float tri15[15], const uint32_t pos[3], float wholeArea, int textured, float outWuv[3])
{
    ClipResult r;
    if (textured) {
        Tri<true> t;
        memcpy(t.v, tri15, sizeof t.v);
        memcpy(t.t, tri15 + 9, sizeof t.t);
        r = clipLeafInVoxel<true>(t, pos[0], pos[1], pos[2], wholeArea);
    }
    else {
        Tri<false> t;
        memcpy(t.v, tri15, sizeof t.v);
        r = clipLeafInVoxel<false>(t, pos[0], pos[1], pos[2], wholeArea);
    }
    outWuv[0] = r.weight;
    outWuv[1] = r.u;
    outWuv[2] = r.v;
    return r.pieces;
}

size_t o2vt_subdivide(const float tri15[15], float *outLeaves, size_t cap)
{
    Tri<true> t;
    memcpy(t.v, tri15, sizeof t.v);
    memcpy(t.t, tri15 + 9, sizeof t.t);
    size_t count = 0;
    forEachLeaf<true>(t, [&](const Tri<true> &leaf) {
        if (count < cap) {
            memcpy(outLeaves + count * 15, leaf.v, sizeof leaf.v);
            memcpy(outLeaves + count * 15 + 9, leaf.t, sizeof leaf.t);
        }
        ++count;
    });
    return count;
}

float o2vt_area(const float v[9])
{
    return triArea(v);
}

int o2vt_aligned(const float v[9])
{
    return triRoughlyAxisAligned(v) ? 1 : 0;
}

void o2vt_transform(const float m[12], const float in[3], float out[3])
{
    affineApply(m, in, out);
}

void o2vt_texture_lookup(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t channels, uint32_t wrap, float u,
                         float v, float rgb[3])
{
    const TextureView tex{pixels, w, h, channels, wrap};
    textureLookup(tex, u, v, rgb);
}

uint32_t o2vt_quantize(const float rgb[3])
{
    return quantizeArgb(rgb[0], rgb[1], rgb[2]);
}

void o2vt_combine(float acc[4], const float incoming[4], int blend)
{
    WeightedColor c{acc[0], acc[1], acc[2], acc[3]};
    combineColorInto(c, incoming[0], incoming[1], incoming[2], incoming[3], blend != 0);
    acc[0] = c.w;
    acc[1] = c.r;
    acc[2] = c.g;
    acc[3] = c.b;
}

}  // extern "C"
