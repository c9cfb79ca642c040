// synth.cu — device side of kasa_b200/synth.py: seeded synthetic genomes and reads,
// so that bench.py can create the C2 workload (1.67 Gbp reference, 100 M reads) directly in HBM.
#include "synth.cuh"

namespace kasa {
namespace {

constexpr uint64_t TAG_GENOME = 1, TAG_TAXON = 2, TAG_OFFSET = 3, TAG_STRAND = 4, TAG_SUBST = 5, TAG_SUBST_TO = 6, TAG_N = 7;

__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t tag) { return splitmix64(seed * 0x100ull + tag); }
__device__ __forceinline__ uint64_t draw(uint64_t skey, uint64_t a, uint64_t b) { return splitmix64(skey ^ ((a << 40) | b)); }

__constant__ char c_code2char[4] = {'A', 'C', 'T', 'G'};   // kASA code order A0 C1 T2 G3

__global__ void __launch_bounds__(256) genome_kernel(uint64_t gkey, uint32_t taxon, uint64_t len, uint32_t marker, char *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len + marker; i += stride)
        out[i] = i < len ? c_code2char[draw(gkey, taxon, i) >> 62] : 'X';
}

__global__ void __launch_bounds__(256) reads_kernel(uint64_t gkey, uint64_t k_tax, uint64_t k_off, uint64_t k_strand, uint64_t k_sub,
                                                    uint64_t k_to, uint64_t k_n, uint64_t first_read, uint32_t n_reads,
                                                    uint32_t n_taxa, uint64_t genome_len, uint32_t read_len, uint32_t subst_ppm,
                                                    uint32_t n_ppm, char *__restrict__ out)
{
    const uint64_t total = (uint64_t)n_reads * read_len;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t r = i / read_len, rid = first_read + r;
        const uint32_t j = (uint32_t)(i % read_len);
        const uint64_t taxon = draw(k_tax, 0, rid) % n_taxa;
        const uint64_t offset = draw(k_off, 0, rid) % (genome_len - read_len + 1);
        const bool rc = draw(k_strand, 0, rid) & 1;
        const uint64_t gpos = rc ? offset + (read_len - 1 - j) : offset + j;
        uint32_t code = (uint32_t)(draw(gkey, taxon, gpos) >> 62);
        if (rc) code ^= 2u;
        const uint64_t bkey = rid * 65536ull + j;
        if (draw(k_sub, 0, bkey) % 1000000ull < subst_ppm) code = (code + (uint32_t)(draw(k_to, 0, bkey) % 3ull) + 1u) & 3u;
        char c = c_code2char[code];
        if (draw(k_n, 0, bkey) % 1000000ull < n_ppm) c = 'N';
        out[i] = c;
    }
}

}  // namespace

void launch_synth_genome(cudaStream_t s, uint64_t seed, uint32_t taxon, uint64_t len, uint32_t marker, char *out, uint64_t *launches)
{
    uint64_t want = div_up<uint64_t>(len + marker, 256);
    unsigned grid = (unsigned)(want < (uint64_t)kSMs * 32 ? want : (uint64_t)kSMs * 32);
    genome_kernel<<<grid, 256, 0, s>>>(stream_key(seed, TAG_GENOME), taxon, len, marker, out);
    ++*launches;
}

void launch_synth_reads(cudaStream_t s, uint64_t genome_seed, uint64_t read_seed, uint64_t first_read, uint32_t n_reads,
                        uint32_t n_taxa, uint64_t genome_len, uint32_t read_len, uint32_t subst_ppm, uint32_t n_ppm, char *out,
                        uint64_t *launches)
{
    if (!n_reads) return;
    uint64_t want = div_up<uint64_t>((uint64_t)n_reads * read_len, 256);
    unsigned grid = (unsigned)(want < (uint64_t)kSMs * 32 ? want : (uint64_t)kSMs * 32);
    reads_kernel<<<grid, 256, 0, s>>>(stream_key(genome_seed, TAG_GENOME), stream_key(read_seed, TAG_TAXON),
                                      stream_key(read_seed, TAG_OFFSET), stream_key(read_seed, TAG_STRAND),
                                      stream_key(read_seed, TAG_SUBST), stream_key(read_seed, TAG_SUBST_TO),
                                      stream_key(read_seed, TAG_N), first_read, n_reads, n_taxa, genome_len, read_len, subst_ppm,
                                      n_ppm, out);
    ++*launches;
}

}  // namespace kasa
