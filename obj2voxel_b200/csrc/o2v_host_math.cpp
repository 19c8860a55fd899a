// Host-side exact arithmetic (compile with -ffp-contract=off): the mesh -> voxel affine transform.
// Restates src/obj2voxel.cpp:370-402 (computeMeshTransform) on top of src/util.hpp:212-281 (AffineTransform ops):
// four affine maps composed left-to-right with row-times-column dot products that start their sum at 0.
#include "o2v_engine.h"

namespace o2v {

namespace {

struct Affine {
    float m[9];
    float t[3];
};

Affine uniformScale(float s, float tx, float ty, float tz)
{
    return Affine{{s, 0, 0, 0, s, 0, 0, 0, s}, {tx, ty, tz}};
}

float dotRowCol(const float *row, float c0, float c1, float c2)
{
    float r = 0;
    r += row[0] * c0;
    r += row[1] * c1;
    r += row[2] * c2;
    return r;
}

/// lhs after rhs: util.hpp:270-281
Affine compose(const Affine &lhs, const Affine &rhs)
{
    Affine out;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            out.m[i * 3 + j] = dotRowCol(lhs.m + i * 3, rhs.m[j], rhs.m[3 + j], rhs.m[6 + j]);
        }
        out.t[i] = dotRowCol(lhs.m + i * 3, rhs.t[0], rhs.t[1], rhs.t[2]);
    }
    for (int i = 0; i < 3; ++i) {
        out.t[i] += lhs.t[i];
    }
    return out;
}

float largest(float a, float b, float c)
{
    const float bc = b < c ? c : b;
    return a < bc ? bc : a;
}

}  // namespace

void computeMeshTransform(const float meshMin[3], const float meshMax[3], uint32_t sampleResolution, const int unit[9],
                          float out[12])
{
    const float antiBleed = 0.5f;
    const float extent = largest(meshMax[0] - meshMin[0], meshMax[1] - meshMin[1], meshMax[2] - meshMin[2]);
    const float sampleScale = static_cast<float>(sampleResolution) - antiBleed;

    Affine acc = uniformScale(1.0f, -meshMin[0], -meshMin[1], -meshMin[2]);          // to the positive octant
    acc = compose(uniformScale(2.0f / extent, -1.0f, -1.0f, -1.0f), acc);           // to [-1, 1]
    Affine axes;
    for (int i = 0; i < 9; ++i) {
        axes.m[i] = static_cast<float>(unit[i]);
    }
    axes.t[0] = axes.t[1] = axes.t[2] = 1.0f;
    acc = compose(axes, acc);                                                        // axis permutation, back to [0, 2]
    acc = compose(uniformScale(sampleScale / 2, antiBleed / 2, antiBleed / 2, antiBleed / 2), acc);  // to the grid

    for (int i = 0; i < 9; ++i) {
        out[i] = acc.m[i];
    }
    for (int i = 0; i < 3; ++i) {
        out[9 + i] = acc.t[i];
    }
}

}  // namespace o2v
