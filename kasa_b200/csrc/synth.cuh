// synth.cuh — device generators matching kasa_b200/synth.py (same splitmix64 streams).
#pragma once
#include "common.cuh"

namespace kasa {

// ASCII genome of `taxon`: len bases followed by `marker` 'X' characters (build marker, Read.hpp:2323-2333)
void launch_synth_genome(cudaStream_t s, uint64_t seed, uint32_t taxon, uint64_t len, uint32_t marker, char *out, uint64_t *launches);
// n_reads x read_len ASCII bases (reads_ascii in synth.py)
void launch_synth_reads(cudaStream_t s, uint64_t genome_seed, uint64_t read_seed, uint64_t first_read, uint32_t n_reads,
                        uint32_t n_taxa, uint64_t genome_len, uint32_t read_len, uint32_t subst_ppm, uint32_t n_ppm, char *out,
                        uint64_t *launches);

}  // namespace kasa
