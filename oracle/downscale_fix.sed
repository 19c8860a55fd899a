# Two-line fix for Voxelizer::downscale() (reference src/voxelization.cpp:538-554, SURVEY fact 3 / Appendix B1):
#   - emplace into the local `result` map instead of the map being iterated/erased
#   - halve each axis of the Morton key (>> 3), not the key itself (/ 2)
s|voxels_.emplace(index / divisor, iter->second)|result.emplace(index >> 3, iter->second)|
