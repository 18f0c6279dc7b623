/* mab_index.h -- host-side index construction and .mai writer (see mab_index.cpp) */
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct MabIdxSeq { std::string name; std::vector<uint8_t> seq; };			/* 1 byte/base codes A,C,G,T,N = 0..4 */
struct MabIdxParams { uint32_t k = 15, w = 10, b = 14, n_frq = 3; float frq[7] = { 0.05f, 0.01f, 0.001f, 0, 0, 0, 0 }; };

/* builds the relocatable index image mab_init() takes (the payload of a .mai block) */
bool mab_build_index(const std::vector<MabIdxSeq> &refs, const MabIdxParams &prm, std::vector<uint8_t> &blob, std::string &err);
/* writes the image as a "PG00" framed .mai file readable by the loader */
bool mab_write_mai(const char *path, const std::vector<uint8_t> &blob);
