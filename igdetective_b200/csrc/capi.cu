// C ABI of libigd_b200.so (declared in include/igd_b200.h): handle management, host-side table
// preparation, H2D/D2H plumbing and launch sequencing.  All computation on contig bases and DP
// cells happens in the CUDA kernels of pack.cu / rss_scan.cu / gotoh.cu; there is no CPU path.
#include "common.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <mutex>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <thread>
#include <unordered_map>

static thread_local std::string g_err;

void igd_set_error(const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
}
int igd_cuda_fail(cudaError_t e, const char *what)
{
    igd_set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return IGD_ERR_CUDA;
}

int igd_reserve(igd_handle *h, int slot, size_t bytes)
{
    if (bytes < 256) bytes = 256;
    if (h->scratch_cap[slot] >= bytes) return IGD_OK;
    if (h->d_scratch[slot]) { IGD_CUDA(cudaFree(h->d_scratch[slot])); h->d_scratch[slot] = nullptr; h->scratch_cap[slot] = 0; }
    const size_t cap = bytes + bytes / 8;
    IGD_CUDA(cudaMalloc(&h->d_scratch[slot], cap));
    h->scratch_cap[slot] = cap;
    return IGD_OK;
}


extern "C" {

const char *igd_last_error(void) { return g_err.c_str(); }

int igd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

igd_handle *igd_create(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        igd_set_error("no CUDA device available (%s); igd_b200 has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n) { igd_set_error("device %d out of range (0..%d)", device, n - 1); return nullptr; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { igd_cuda_fail(e, "cudaGetDeviceProperties"); return nullptr; }
    if (prop.major != 10) {
        igd_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { igd_cuda_fail(e, "cudaSetDevice"); return nullptr; }
    igd_handle *h = new igd_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess
           && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess && cudaEventCreate(&h->ev2) == cudaSuccess
           && cudaMalloc(&h->d_count, 64 * sizeof(unsigned long long)) == cudaSuccess
           && cudaMalloc(&h->d_lut7, 16384) == cudaSuccess && cudaMalloc(&h->d_lut9, 262144) == cudaSuccess
           && cudaMalloc(&h->d_any9, 32768) == cudaSuccess && cudaMalloc(&h->d_any9c, 32768) == cudaSuccess && cudaMalloc(&h->d_need7, 16384) == cudaSuccess && cudaMalloc(&h->d_any7, 2048) == cudaSuccess;
    if (!ok) { igd_cuda_fail(cudaGetLastError(), "igd_create allocations"); igd_destroy(h); return nullptr; }
    if (igd_probe_fp_increment(h, &h->fp_increment_ok) != IGD_OK) { igd_destroy(h); return nullptr; }
    return h;
}

void igd_destroy(igd_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_comp); cudaFree(h->d_lut7); cudaFree(h->d_lut9); cudaFree(h->d_any9); cudaFree(h->d_any9c); cudaFree(h->d_need7); cudaFree(h->d_any7);
    cudaFree(h->d_planes); cudaFree(h->d_inv); cudaFree(h->d_invsum);
    cudaFree(h->d_hits); cudaFree(h->d_count);
    for (int i = 0; i < S_COUNT; ++i) cudaFree(h->d_scratch[i]);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    for (int i = 0; i < 3; ++i) if (h->h_pin[i]) cudaFreeHost(h->h_pin[i]);
    if (h->h_ring) cudaFreeHost(h->h_ring);
    for (int i = 0; i < 4; ++i) if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int igd_sync(igd_handle *h)
{
    if (!h) { igd_set_error("null handle"); return IGD_ERR_ARG; }
    IGD_CUDA(cudaSetDevice(h->device));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    return IGD_OK;
}

int igd_get_timing(igd_handle *h, igd_timing *out)
{
    if (!h || !out) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    *out = h->timing;
    return IGD_OK;
}

// ---- motifs ---------------------------------------------------------------------------------------

// natural index (first letter most significant, A0 C1 G2 T3) -> kernel index ((hi bits << k) | lo bits)
static inline unsigned kernel_index(unsigned nat, int k)
{
    unsigned lo = 0, hi = 0;
    for (int i = 0; i < k; ++i) {
        const unsigned code = (nat >> (2 * (k - 1 - i))) & 3u;
        lo |= (code & 1u) << i; hi |= (code >> 1) << i;
    }
    return (hi << k) | lo;
}
static inline unsigned revcomp_index(unsigned nat, int k)
{
    unsigned r = 0;
    for (int i = 0; i < k; ++i) { r = (r << 2) | (3u - (nat & 3u)); nat >>= 2; }
    return r;
}

// Centre bitmap of the RSS scan: bit c (kernel index order) is set iff the 9-mer c, or a 9-mer that
// overlaps it shifted by one base to the left or to the right, is in some nonamer set on either strand.
// The three candidate nonamers of a spacer window (s-1, s, s+1; py/IGDetective.py:162-168) start at
// x, x+1, x+2, so the 9-mer at x+1 missing from this bitmap proves that none of the three is a motif.
int igd_center_bitmap(const uint8_t *flags9, uint32_t *bitmap)
{
    if (!flags9 || !bitmap) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    memset(bitmap, 0, 32768);
    auto set = [&](unsigned nat) { const unsigned ki = kernel_index(nat & 0x3ffffu, 9); bitmap[ki >> 5] |= 1u << (ki & 31); };
    for (unsigned nat = 0; nat < 262144u; ++nat) {
        if (!flags9[nat] && !flags9[revcomp_index(nat, 9)]) continue;
        set(nat);
        for (unsigned a = 0; a < 4; ++a) {
            set((nat << 2) | a);                 // member starts one base before the centre
            set((nat >> 2) | (a << 16));         // member starts one base after the centre
        }
    }
    return IGD_OK;
}

int igd_set_motifs(igd_handle *h, const uint8_t *flags7, const uint8_t *flags9)
{
    if (!h || !flags7 || !flags9) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    IGD_CUDA(cudaSetDevice(h->device));
    std::vector<uint8_t> l7(16384), l9(262144);
    std::vector<uint32_t> a7(512, 0), a9(8192, 0);
    for (int k = 7; k <= 9; k += 2) {
        const unsigned n = 1u << (2 * k);
        const uint8_t *src = (k == 7) ? flags7 : flags9;
        std::vector<uint8_t> &dst = (k == 7) ? l7 : l9;
        std::vector<uint32_t> &any = (k == 7) ? a7 : a9;
        for (unsigned nat = 0; nat < n; ++nat) {
            if (src[nat] & 0xF0) { igd_set_error("motif flags use bits 0..3 only"); return IGD_ERR_ARG; }
            const uint8_t f = (uint8_t)(src[nat] | (src[revcomp_index(nat, k)] << 4));
            const unsigned ki = kernel_index(nat, k);
            dst[ki] = f;
            if (f) any[ki >> 5] |= 1u << (ki & 31);
        }
    }
    // does the bit-sliced consensus prefilter of rss_scan.cu cover every heptamer of these tables?
    bool covered = true;
    for (unsigned nat = 0; nat < 16384u && covered; ++nat) {
        if (!flags7[nat] && !flags7[revcomp_index(nat, 7)]) continue;
        static const int want[7] = {1, 0, 1, -1, 2, 3, 2};           // C A C . G T G
        int m = 0;
        for (int i = 0; i < 7; ++i) m += (int)((nat >> (2 * (6 - i))) & 3u) == want[i];
        covered = m >= 4;
    }
    h->consensus_prefilter = covered;
    IGD_CUDA(cudaMemcpyAsync(h->d_lut7, l7.data(), 16384, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_lut9, l9.data(), 262144, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_any7, a7.data(), 2048, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_any9, a9.data(), 32768, cudaMemcpyHostToDevice, h->stream));
    std::vector<uint32_t> a9c(8192);
    igd_center_bitmap(flags9, a9c.data());
    IGD_CUDA(cudaMemcpyAsync(h->d_any9c, a9c.data(), 32768, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    h->h_lut7 = l7;
    h->need7_valid = false;
    h->motifs_set = true;
    return IGD_OK;
}

// ---- contigs --------------------------------------------------------------------------------------

// Letters -> device.  Pinned (or registered) caller memory goes in one DMA.  Pageable memory of a whole genome
// would cross the bus at the driver's staging rate (~11 GB/s measured); instead a few host threads, started once per
// upload, each own one pinned slot of a small ring (4 x 8 MiB: pinning memory is itself slow, and several ranks of
// one box do it at once) and walk over every fourth piece: copy the piece into the slot, queue its DMA, record the
// slot's event, and wait for that event before the slot is written again -- so the memcpy of one thread overlaps the
// transfers of the others.  The copies land in disjoint ranges; whatever follows on the stream is queued after the join.
static int upload_letters(igd_handle *h, uint8_t *d_dst, const uint8_t *src, size_t nbytes, bool throttle = false)
{
    constexpr size_t PIECE = 8u << 20;
    constexpr int RING = 4;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();                                    // unregistered host memory may report an error on old drivers
    if (pinned && !throttle) {
        IGD_CUDA(cudaMemcpyAsync(d_dst, src, nbytes, cudaMemcpyHostToDevice, h->stream));
        return IGD_OK;
    }
    if (pinned) {
        // A copy engine works through its queue in order: behind one 0.8 GB transfer -- or behind a hundred queued pieces
        // of it -- the small table copies of another handle of this GPU (a pipeline that scores one set of contigs while
        // the next is uploaded) wait for milliseconds each.  So: pieces, and never more than two of them queued.
        if (!h->ring_ev[0])
            for (int i = 0; i < RING; ++i) IGD_CUDA(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
        size_t k = 0;
        for (size_t at = 0; at < nbytes; at += PIECE, ++k) {
            if (k >= 2) IGD_CUDA(cudaEventSynchronize(h->ring_ev[k & 1]));
            IGD_CUDA(cudaMemcpyAsync(d_dst + at, src + at, std::min(PIECE, nbytes - at), cudaMemcpyHostToDevice, h->stream));
            IGD_CUDA(cudaEventRecord(h->ring_ev[k & 1], h->stream));
        }
        return IGD_OK;
    }
    if (nbytes < (64u << 20)) {
        IGD_CUDA(cudaMemcpyAsync(d_dst, src, nbytes, cudaMemcpyHostToDevice, h->stream));
        return IGD_OK;
    }
    if (!h->h_ring) {
        IGD_CUDA(cudaMallocHost(&h->h_ring, RING * PIECE));
        if (!h->ring_ev[0])
            for (int i = 0; i < RING; ++i) IGD_CUDA(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
    }
    const size_t n_pieces = (nbytes + PIECE - 1) / PIECE;
    const unsigned nthreads = std::max(1u, std::min((unsigned)RING, std::thread::hardware_concurrency()));
    cudaError_t errs[RING];
    for (int i = 0; i < RING; ++i) errs[i] = cudaSuccess;
    auto worker = [&](unsigned t) {
        cudaError_t e = cudaSetDevice(h->device);
        uint8_t *slot = (uint8_t *)h->h_ring + (size_t)t * PIECE;
        bool used = false;
        for (size_t i = t; i < n_pieces && e == cudaSuccess; i += nthreads) {
            const size_t at = i * PIECE, n = std::min(PIECE, nbytes - at);
            if (used) e = cudaEventSynchronize(h->ring_ev[t]);          // the DMA that last read this slot
            if (e != cudaSuccess) break;
            memcpy(slot, src + at, n);
            e = cudaMemcpyAsync(d_dst + at, slot, n, cudaMemcpyHostToDevice, h->stream);
            if (e == cudaSuccess) e = cudaEventRecord(h->ring_ev[t], h->stream);
            used = true;
        }
        errs[t] = e;
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &th : pool) th.join();
    for (unsigned t = 0; t < nthreads; ++t) IGD_CUDA(errs[t]);
    return IGD_OK;
}

} // extern "C"

// wait = false: returns once the caller's letters have crossed the bus (they may be reused); the pack kernel stays queued
// on the handle's stream, where every later call of the handle is ordered behind it
static int load_contigs_impl(igd_handle *h, const uint8_t *seq, const uint64_t *offsets, int32_t n_contigs, bool wait)
{
    if (!h || !offsets || n_contigs < 0 || (!seq && n_contigs > 0 && offsets[n_contigs] > 0)) {
        igd_set_error("bad arguments to igd_load_contigs"); return IGD_ERR_ARG;
    }
    IGD_CUDA(cudaSetDevice(h->device));
    h->contigs_loaded = false;
    h->contig_chunk0.assign(n_contigs, 0);
    h->contig_len.assign(n_contigs, 0);
    uint64_t chunk = 1;                               // chunk 0 is the leading pad
    for (int c = 0; c < n_contigs; ++c) {
        if (offsets[c + 1] < offsets[c]) { igd_set_error("offsets must be non-decreasing"); return IGD_ERR_ARG; }
        const uint64_t len = offsets[c + 1] - offsets[c];
        h->contig_chunk0[c] = chunk;
        h->contig_len[c] = len;
        chunk += (len + 63) / 64 + 1;                 // data chunks + one pad chunk
    }
    h->total_bases = n_contigs ? offsets[n_contigs] - offsets[0] : 0;
    const uint64_t data_chunks = chunk - 1;
    h->n_tiles = (data_chunks + IGD_TILE_CHUNKS - 1) / IGD_TILE_CHUNKS;
    if (h->n_tiles == 0) h->n_tiles = 1;
    h->n_chunks = h->n_tiles * IGD_TILE_CHUNKS + 2;
    if (h->n_chunks > h->planes_cap) {
        cudaFree(h->d_planes); cudaFree(h->d_inv); cudaFree(h->d_invsum);
        h->d_planes = nullptr; h->d_inv = nullptr; h->d_invsum = nullptr; h->planes_cap = 0;
        const uint64_t cap = h->n_chunks + h->n_chunks / 16;
        IGD_CUDA(cudaMalloc(&h->d_planes, cap * sizeof(uint4)));
        IGD_CUDA(cudaMalloc(&h->d_inv, cap * sizeof(unsigned long long)));
        IGD_CUDA(cudaMalloc(&h->d_invsum, ((cap + 31) / 32 + 1) * sizeof(uint32_t)));
        h->planes_cap = cap;
    }
    const uint64_t nbytes = h->total_bases;
    int rc;
    if ((rc = igd_reserve(h, S_SEQ, nbytes))) return rc;
    if ((rc = igd_reserve(h, S_SEQOFF, (size_t)(n_contigs + 1) * 8))) return rc;
    if ((rc = igd_reserve(h, S_CHUNK0, (size_t)(n_contigs + 1) * 8))) return rc;
    std::vector<uint64_t> rel(n_contigs + 1);
    for (int c = 0; c <= n_contigs; ++c) rel[c] = offsets[c] - offsets[0];
    h->contig_seq_off = rel;
    if (nbytes) {
        int rc2 = upload_letters(h, (uint8_t *)h->d_scratch[S_SEQ], seq + offsets[0], nbytes, !wait);
        if (rc2) return rc2;
    }
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_SEQOFF], rel.data(), (size_t)(n_contigs + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    if (n_contigs)
        IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_CHUNK0], h->contig_chunk0.data(), (size_t)n_contigs * 8, cudaMemcpyHostToDevice, h->stream));
    // contig holding the first chunk of every tile of 512 chunks (-1: the leading pad), for the pack kernel
    std::vector<int32_t> tile_contig((size_t)h->n_tiles + 1);
    {
        int c = -1;
        for (uint64_t t = 0; t <= h->n_tiles; ++t) {
            const uint64_t ch = t * IGD_TILE_CHUNKS;
            while (c + 1 < n_contigs && h->contig_chunk0[c + 1] <= ch) ++c;
            tile_contig[t] = c;
        }
    }
    if ((rc = igd_reserve(h, S_TILEC, tile_contig.size() * 4))) return rc;
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TILEC], tile_contig.data(), tile_contig.size() * 4, cudaMemcpyHostToDevice, h->stream));
    // hit buffer sized for the k-mer scan's ~1 % density with slack; regrown on overflow
    const uint64_t want = std::max<uint64_t>(1u << 20, h->total_bases / 32);
    if (want > h->hit_cap) {
        cudaFree(h->d_hits); h->d_hits = nullptr; h->hit_cap = 0;
        IGD_CUDA(cudaMalloc(&h->d_hits, want * sizeof(unsigned long long)));
        h->hit_cap = want;
    }
    IGD_CUDA(cudaEventRecord(h->ev2, h->stream));     // every copy from host memory is in front of this
    IGD_CUDA(cudaEventRecord(h->ev0, h->stream));
    if ((rc = igd_launch_pack(h, (const uint8_t *)h->d_scratch[S_SEQ], (const unsigned long long *)h->d_scratch[S_SEQOFF],
                              (const unsigned long long *)h->d_scratch[S_CHUNK0], (const int *)h->d_scratch[S_TILEC], n_contigs))) return rc;
    IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
    if (wait) {
        IGD_CUDA(cudaStreamSynchronize(h->stream));
        IGD_CUDA(cudaEventElapsedTime(&h->timing.pack_ms, h->ev0, h->ev1));
    } else {
        IGD_CUDA(cudaEventSynchronize(h->ev2));       // rel / offsets / the caller's letters are free again
        h->timing.pack_ms = 0.f;                      // not waited for
    }
    h->contigs_loaded = true;
    h->last_scan_kind = 0;
    h->sorted_hits_valid = false;
    return IGD_OK;
}

extern "C" {

int igd_load_contigs(igd_handle *h, const uint8_t *seq, const uint64_t *offsets, int32_t n_contigs)
{
    return load_contigs_impl(h, seq, offsets, n_contigs, true);
}

int igd_load_contigs_async(igd_handle *h, const uint8_t *seq, const uint64_t *offsets, int32_t n_contigs)
{
    return load_contigs_impl(h, seq, offsets, n_contigs, false);
}

// ---- FASTA ingest -----------------------------------------------------------------------------------
} // extern "C"

// Parses the lines of the file whose first byte lies in [lo, hi), plus `margin` bytes of lines on either side when
// `ranged` (sharded runs: the margins are the context a neighbouring range's records need, igd_fasta_open_range).
// Parallel parse: the span is cut at line starts into one piece per thread; every piece is parsed into its own buffer
// (letters with blanks removed + the headers it contains), then the pieces are concatenated by a parallel copy.
// Whole-file mode: letters in front of the file's first header are dropped and ids follow the reference's dict
// semantics.  Range mode: record 0 is the lead (the letters in front of the first header: they continue a record of an
// earlier range), records keep file order, nothing is merged.
static igd_fasta *fasta_parse(const char *path, uint64_t lo_req, uint64_t hi_req, uint64_t margin, bool ranged)
{
    if (!path) { igd_set_error("null path"); return nullptr; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { igd_set_error("cannot open %s", path); return nullptr; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); igd_set_error("cannot stat %s", path); return nullptr; }
    const size_t fsize = (size_t)st.st_size;
    const char *data = fsize ? (const char *)mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (fsize && data == MAP_FAILED) { igd_set_error("cannot map %s", path); return nullptr; }
    if (fsize >= 2 && (unsigned char)data[0] == 0x1f && (unsigned char)data[1] == 0x8b) {
        munmap((void *)data, fsize);
        igd_set_error("%s is gzip-compressed; igd_fasta_open reads plain FASTA", path); return nullptr;
    }
    if (fsize && madvise((void *)data, fsize, MADV_SEQUENTIAL) != 0) { /* advisory only */ }
    // first line start at or after byte p (p itself when it is 0 or follows a newline)
    auto line_start = [&](size_t p) -> size_t {
        if (p == 0) return 0;
        if (p >= fsize) return fsize;
        const char *nl = (const char *)memchr(data + p - 1, '\n', fsize - (p - 1));
        return nl ? (size_t)(nl - data) + 1 : fsize;
    };
    const size_t own_lo = ranged ? line_start((size_t)std::min<uint64_t>(lo_req, fsize)) : 0;
    const size_t own_hi = ranged ? std::max(own_lo, line_start((size_t)std::min<uint64_t>(hi_req, fsize))) : fsize;
    const size_t ext_lo = ranged ? line_start(own_lo > margin ? own_lo - (size_t)margin : 0) : 0;
    const size_t ext_hi = ranged ? line_start((size_t)std::min<uint64_t>((uint64_t)own_hi + margin, fsize)) : fsize;
    struct Hdr { std::string id; uint64_t at; uint64_t byte; };     // at = letters of the piece in front of the header
    struct Piece { uint8_t *buf = nullptr; uint64_t n = 0; std::vector<Hdr> hdrs; };
    const size_t own_bytes = own_hi - own_lo;
    unsigned nthreads = std::max(1u, std::min(std::min(std::thread::hardware_concurrency(), 32u), (unsigned)(own_bytes / (32u << 20)) + 1u));
    // pieces: [pre-margin] own span in nthreads pieces [post-margin]; empty pieces are harmless
    std::vector<size_t> cut;
    cut.push_back(ext_lo);
    cut.push_back(own_lo);
    for (unsigned k = 1; k < nthreads; ++k) {
        size_t p = std::max(cut.back(), own_lo + own_bytes / nthreads * k);
        const char *nl = p < own_hi ? (const char *)memchr(data + p, '\n', own_hi - p) : nullptr;
        cut.push_back(nl ? std::min((size_t)(nl - data) + 1, own_hi) : own_hi);
    }
    cut.push_back(own_hi);
    cut.push_back(ext_hi);
    const unsigned npieces = (unsigned)cut.size() - 1;
    std::vector<Piece> pieces(npieces);
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; };
    auto parse = [&](unsigned k) {
        Piece &pc = pieces[k];
        const char *p = data + cut[k], *end = data + cut[k + 1];
        pc.buf = (uint8_t *)malloc((size_t)(end - p) + 1);
        uint8_t *w = pc.buf;
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *le = nl ? nl : end;             // line = [p, le)
            if (p < le && *p == '>') {
                const char *q = p + 1;
                while (q < le && blank(*q)) ++q;         // the id is the first token of the header
                const char *b = q;
                while (q < le && !blank(*q)) ++q;
                pc.hdrs.push_back({std::string(b, q), (uint64_t)(w - pc.buf), (uint64_t)(p - data)});
            } else {
                const char *q = p;
                while (q < le) {                        // copy the runs between blanks
                    const char *s = q;
                    while (q < le && !blank(*q)) ++q;
                    memcpy(w, s, (size_t)(q - s)); w += q - s;
                    while (q < le && blank(*q)) ++q;
                }
            }
            p = nl ? nl + 1 : end;
        }
        pc.n = (uint64_t)(w - pc.buf);
    };
    {
        std::vector<std::thread> th;
        for (unsigned k = 1; k < npieces; ++k) th.emplace_back(parse, k);
        parse(0);
        for (auto &t : th) t.join();
    }
    if (fsize) munmap((void *)data, fsize);
    // global letter offset of every piece; whole-file mode drops everything in front of the first header
    unsigned first_piece = npieces;
    for (unsigned k = 0; k < npieces && first_piece == npieces; ++k) if (!pieces[k].hdrs.empty()) first_piece = k;
    std::vector<uint64_t> base(npieces + 1, 0), skip(npieces, 0);
    for (unsigned k = 0; k < npieces; ++k) {
        if (!ranged) skip[k] = k < first_piece ? pieces[k].n : (k == first_piece ? pieces[k].hdrs[0].at : 0);
        base[k + 1] = base[k] + pieces[k].n - skip[k];
    }
    const uint64_t total = base[npieces];
    struct Rec { std::string id; uint64_t off, len, byte; };
    std::vector<Rec> recs;
    if (ranged) recs.push_back({std::string(), 0, 0, UINT64_MAX});          // the lead
    for (unsigned k = ranged ? 0 : first_piece; k < npieces; ++k)
        for (const Hdr &hd : pieces[k].hdrs) recs.push_back({hd.id, base[k] + hd.at - skip[k], 0, hd.byte});
    for (size_t i = 0; i < recs.size(); ++i) recs[i].len = (i + 1 < recs.size() ? recs[i + 1].off : total) - recs[i].off;
    uint8_t *all = (uint8_t *)malloc((size_t)total + 1);
    {
        auto copy = [&](unsigned k) { memcpy(all + base[k], pieces[k].buf + skip[k], (size_t)(pieces[k].n - skip[k])); free(pieces[k].buf); };
        std::vector<std::thread> th;
        for (unsigned k = 1; k < npieces; ++k) th.emplace_back(copy, k);
        copy(0);
        for (auto &t : th) t.join();
    }
    igd_fasta *f = new igd_fasta();
    f->range_lo = own_lo; f->range_hi = own_hi;
    f->own_lo = base[1]; f->own_hi = base[npieces - 1];
    if (ranged) {
        f->seq = all;
        f->off.assign(1, 0);
        for (const Rec &r : recs) { f->ids.push_back(r.id); f->off.push_back(r.off + r.len); f->hdr_byte.push_back(r.byte); }
        return f;
    }
    // dict semantics: a repeated id keeps its first position and takes the last sequence
    std::unordered_map<std::string, size_t> first;
    std::vector<size_t> order;
    for (size_t i = 0; i < recs.size(); ++i) {
        auto it = first.find(recs[i].id);
        if (it == first.end()) { first.emplace(recs[i].id, order.size()); order.push_back(i); }
        else order[it->second] = i;
    }
    f->off.assign(1, 0);
    const bool in_place = order.size() == recs.size();
    for (size_t k = 0; k < order.size(); ++k) {
        const Rec &r = recs[order[k]];
        f->ids.push_back(r.id);
        f->off.push_back(f->off.back() + r.len);
        f->hdr_byte.push_back(r.byte);
    }
    if (in_place) {
        f->seq = all;
    } else {                                            // rare: duplicates -> rebuild in dict order
        f->seq = (uint8_t *)malloc((size_t)f->off.back() + 1);
        for (size_t k = 0; k < order.size(); ++k)
            memcpy(f->seq + f->off[k], all + recs[order[k]].off, (size_t)recs[order[k]].len);
        free(all);
    }
    return f;
}

extern "C" {

igd_fasta *igd_fasta_open(const char *path) { return fasta_parse(path, 0, UINT64_MAX, 0, false); }

igd_fasta *igd_fasta_open_range(const char *path, uint64_t byte_lo, uint64_t byte_hi, uint64_t margin_bytes)
{
    if (byte_hi < byte_lo) { igd_set_error("byte_hi < byte_lo"); return nullptr; }
    return fasta_parse(path, byte_lo, byte_hi, margin_bytes, true);
}

int igd_fasta_range(const igd_fasta *f, uint64_t *byte_lo, uint64_t *byte_hi, uint64_t *own_lo_letter, uint64_t *own_hi_letter)
{
    if (!f) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    if (byte_lo) *byte_lo = f->range_lo;
    if (byte_hi) *byte_hi = f->range_hi;
    if (own_lo_letter) *own_lo_letter = f->own_lo;
    if (own_hi_letter) *own_hi_letter = f->own_hi;
    return IGD_OK;
}

int igd_fasta_header_byte(const igd_fasta *f, int32_t i, uint64_t *byte)
{
    if (!f || !byte || i < 0 || (size_t)i >= f->ids.size()) { igd_set_error("contig index out of range"); return IGD_ERR_ARG; }
    *byte = f->hdr_byte[(size_t)i];
    return IGD_OK;
}

void igd_fasta_close(igd_fasta *f) { delete f; }

int32_t igd_fasta_count(const igd_fasta *f) { return f ? (int32_t)f->ids.size() : 0; }

int igd_fasta_info(const igd_fasta *f, int32_t i, const char **id, uint64_t *length)
{
    if (!f || i < 0 || (size_t)i >= f->ids.size()) { igd_set_error("contig index out of range"); return IGD_ERR_ARG; }
    if (id) *id = f->ids[i].c_str();
    if (length) *length = f->off[i + 1] - f->off[i];
    return IGD_OK;
}

int igd_fasta_read(const igd_fasta *f, int32_t i, uint64_t start, uint64_t len, uint8_t *out)
{
    if (!f || i < 0 || (size_t)i >= f->ids.size() || (!out && len)) { igd_set_error("bad arguments to igd_fasta_read"); return IGD_ERR_ARG; }
    const uint64_t L = f->off[i + 1] - f->off[i];
    if (start > L || len > L - start) {
        igd_set_error("slice [%llu, +%llu) outside contig of %llu", (unsigned long long)start, (unsigned long long)len, (unsigned long long)L);
        return IGD_ERR_ARG;
    }
    memcpy(out, f->seq + f->off[i] + start, (size_t)len);
    return IGD_OK;
}

int igd_load_fasta(igd_handle *h, const igd_fasta *f)
{
    if (!h || !f) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    return igd_load_contigs(h, f->seq, f->off.data(), (int32_t)f->ids.size());
}

static int run_scan(igd_handle *h, int kind, int k, const int32_t *spacer, uint64_t *n_hits)
{
    if (!h || !n_hits) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    if (!h->motifs_set || !h->contigs_loaded) { igd_set_error("igd_set_motifs and igd_load_contigs must precede a scan"); return IGD_ERR_STATE; }
    IGD_CUDA(cudaSetDevice(h->device));
    h->timing.scan_launches = 0;
    if (kind == 1) { int rc0 = igd_launch_rss_scan(h, spacer, true); if (rc0) return rc0; }
    for (int attempt = 0; attempt < 2; ++attempt) {
        IGD_CUDA(cudaMemsetAsync(h->d_count, 0, 2 * sizeof(unsigned long long), h->stream));   // hits, strip counter
        IGD_CUDA(cudaEventRecord(h->ev0, h->stream));
        int rc = (kind == 1) ? igd_launch_rss_scan(h, spacer) : igd_launch_kmer_scan(h, k);
        if (rc) return rc;
        h->timing.scan_launches++;
        IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
        unsigned long long count = 0;
        IGD_CUDA(cudaMemcpyAsync(&count, h->d_count, sizeof count, cudaMemcpyDeviceToHost, h->stream));
        IGD_CUDA(cudaStreamSynchronize(h->stream));
        IGD_CUDA(cudaEventElapsedTime(&h->timing.scan_ms, h->ev0, h->ev1));
        h->timing.scan_bases = h->total_bases;
        if (count <= h->hit_cap) {
            *n_hits = count;
            h->host_hits.clear();
            h->sorted_hits_valid = false;
            h->last_scan_kind = kind; h->last_k = k;
            if (spacer) memcpy(h->last_spacer, spacer, sizeof h->last_spacer);
            return IGD_OK;
        }
        // overflow: grow to the observed count and rerun once
        cudaFree(h->d_hits); h->d_hits = nullptr; h->hit_cap = 0;
        const uint64_t want = count + count / 8 + 1024;
        IGD_CUDA(cudaMalloc(&h->d_hits, want * sizeof(unsigned long long)));
        h->hit_cap = want;
    }
    igd_set_error("hit buffer overflow persisted after regrow");
    return IGD_ERR_CAPACITY;
}

} // extern "C"

int igd_run_rss_scan(igd_handle *h, const int32_t spacer[4], uint64_t *n_hits)
{
    if (!spacer) { igd_set_error("null spacer"); return IGD_ERR_ARG; }
    for (int t = 0; t < 4; ++t)
        if (spacer[t] < 2 || spacer[t] > 23) {
            // the staged halo is one chunk (64 bases): 9 + (s+1) <= 33 and 7 + (s+1) + 9 <= 40 must hold
            igd_set_error("spacer[%d]=%d outside the supported range 2..23", t, spacer[t]); return IGD_ERR_ARG;
        }
    return run_scan(h, 1, 0, spacer, n_hits);
}

extern "C" {

int igd_scan(igd_handle *h, const int32_t spacer[4], uint64_t *n_hits) { return igd_run_rss_scan(h, spacer, n_hits); }

int igd_scan_kmers(igd_handle *h, int32_t k, uint64_t *n_hits)
{
    if (k != 7 && k != 9) { igd_set_error("k must be 7 or 9"); return IGD_ERR_ARG; }
    return run_scan(h, 2, k, nullptr, n_hits);
}

static int pull_hits(igd_handle *h, uint64_t n)
{
    if (h->host_hits.size() == n) return IGD_OK;
    h->host_hits.resize(n);
    if (n) {
        IGD_CUDA(cudaMemcpyAsync(h->host_hits.data(), h->d_hits, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        IGD_CUDA(cudaStreamSynchronize(h->stream));
    }
    return IGD_OK;
}

// contig holding chunk index ch
static inline uint32_t contig_of(const igd_handle *h, uint64_t ch)
{
    return (uint32_t)(std::upper_bound(h->contig_chunk0.begin(), h->contig_chunk0.end(), ch) - h->contig_chunk0.begin() - 1);
}

static int current_count(igd_handle *h, uint64_t *n)
{
    unsigned long long count = 0;
    IGD_CUDA(cudaMemcpyAsync(&count, h->d_count, sizeof count, cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    *n = count;
    return IGD_OK;
}

} // extern "C"

// Decodes the raw keys of the last RSS scan into igd_hit records sorted by (contig, sig_type, strand, index of the 5'
// k-mer in scanned-strand coordinates); cached in the handle until the next scan.
int igd_sorted_hits(igd_handle *h, uint64_t *n_out)
{
    if (h->last_scan_kind != 1) { igd_set_error("no RSS scan result to fetch"); return IGD_ERR_STATE; }
    if (h->sorted_hits_valid) { *n_out = h->sorted_hits.size(); return IGD_OK; }
    IGD_CUDA(cudaSetDevice(h->device));
    uint64_t n = 0; int rc;
    if ((rc = current_count(h, &n))) return rc;
    const bool on_device = !(getenv("IGD_HITS_ON_HOST") && atoi(getenv("IGD_HITS_ON_HOST")) != 0);      // read per call: tests compare both
    if (on_device) {
        bool done = false;
        if ((rc = igd_sort_hits_device(h, n, &done))) return rc;
        if (done) { h->sorted_hits_valid = true; *n_out = n; return IGD_OK; }
    }
    if ((rc = pull_hits(h, n))) return rc;
    // Host path (coordinates beyond the device's order key, or IGD_HITS_ON_HOST=1): one pass decodes every key (contig by binary search over the chunk table) and builds a 64-bit sort key
    // contig (26 bits) | type (2) | strand (1) | index of the list's first element in scanned-strand coordinates (35);
    // one sort of 16-byte (key, index) pairs gives the final order.
    std::vector<igd_hit> hits(n);
    std::vector<std::pair<uint64_t, uint32_t>> keyed(n);
    bool fits = h->contig_chunk0.size() < (1u << 26) && n < 0xffffffffull;
    for (uint64_t i = 0; i < n; ++i) {
        const unsigned long long key = h->host_hits[i];
        const uint64_t g = key >> 8;
        const int t = (int)((key >> 4) & 3), strand = (int)((key >> 3) & 1), d = (int)(key & 3);
        const uint32_t c = contig_of(h, g >> 6);
        const int64_t hept = (int64_t)(g - h->contig_chunk0[c] * 64);
        const int s1 = h->last_spacer[t] + (d == 0 ? 0 : (d == 1 ? -1 : 1));
        const bool hept_first = (t == IGD_SIG_V || t == IGD_SIG_D_RIGHT);
        int64_t nona;
        if (strand == 0) nona = hept_first ? hept + 7 + s1 : hept - 9 - s1;
        else             nona = hept_first ? hept - s1 - 9 : hept + 7 + s1;
        const int64_t L = (int64_t)h->contig_len[c];
        const int64_t first = hept_first ? hept : nona;
        const int64_t kf = hept_first ? 7 : 9;
        const int64_t scan_idx = strand == 0 ? first : L - first - kf;
        igd_hit &r = hits[i];
        r.hept = hept; r.nona = nona; r.contig = c;
        r.sig_type = (uint8_t)t; r.strand = (uint8_t)strand; r.spacer = (uint8_t)s1; r.reserved = 0;
        fits = fits && scan_idx >= 0 && scan_idx < (1ll << 35);
        keyed[i] = {((uint64_t)c << 38) | ((uint64_t)t << 36) | ((uint64_t)strand << 35) | (uint64_t)scan_idx, (uint32_t)i};
    }
    if (fits) {
        std::sort(keyed.begin(), keyed.end());
    } else {                                        // contigs beyond 2^35 bases or 2^26 contigs: compare field by field
        std::sort(keyed.begin(), keyed.end(), [&](const std::pair<uint64_t, uint32_t> &x, const std::pair<uint64_t, uint32_t> &y) {
            const igd_hit &a = hits[x.second], &b = hits[y.second];
            if (a.contig != b.contig) return a.contig < b.contig;
            if (a.sig_type != b.sig_type) return a.sig_type < b.sig_type;
            if (a.strand != b.strand) return a.strand < b.strand;
            const bool hf = (a.sig_type == IGD_SIG_V || a.sig_type == IGD_SIG_D_RIGHT);
            const int64_t fa = hf ? a.hept : a.nona, fb = hf ? b.hept : b.nona;
            return a.strand == 0 ? fa < fb : fa > fb;
        });
    }
    h->sorted_hits.resize(n);
    for (uint64_t i = 0; i < n; ++i) h->sorted_hits[i] = hits[keyed[i].second];
    h->sorted_hits_valid = true;
    *n_out = n;
    return IGD_OK;
}

extern "C" {

int igd_fetch_hits(igd_handle *h, igd_hit *out, uint64_t capacity)
{
    if (!h || (!out && capacity)) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    uint64_t n = 0; int rc;
    if ((rc = igd_sorted_hits(h, &n))) return rc;
    if (n > capacity) { igd_set_error("capacity %llu < %llu hits", (unsigned long long)capacity, (unsigned long long)n); return IGD_ERR_CAPACITY; }
    if (n) memcpy(out, h->sorted_hits.data(), n * sizeof(igd_hit));
    return IGD_OK;
}

int igd_fetch_kmer_hits(igd_handle *h, igd_kmer_hit *out, uint64_t capacity)
{
    if (!h || (!out && capacity)) { igd_set_error("null argument"); return IGD_ERR_ARG; }
    if (h->last_scan_kind != 2) { igd_set_error("no k-mer scan result to fetch"); return IGD_ERR_STATE; }
    IGD_CUDA(cudaSetDevice(h->device));
    uint64_t n = 0; int rc;
    if ((rc = current_count(h, &n))) return rc;
    if (n > capacity) { igd_set_error("capacity %llu < %llu hits", (unsigned long long)capacity, (unsigned long long)n); return IGD_ERR_CAPACITY; }
    if ((rc = pull_hits(h, n))) return rc;
    std::sort(h->host_hits.begin(), h->host_hits.end());
    for (uint64_t i = 0; i < n; ++i) {
        const unsigned long long key = h->host_hits[i];
        const uint64_t g = key >> 8;
        const uint32_t c = contig_of(h, g >> 6);
        out[i].pos = (int64_t)(g - h->contig_chunk0[c] * 64);
        out[i].contig = c;
        out[i].flags = (uint8_t)(key & 0xff);
        out[i].reserved[0] = out[i].reserved[1] = out[i].reserved[2] = 0;
    }
    return IGD_OK;
}

// ---- alignment ------------------------------------------------------------------------------------

// Can every intermediate DP value of an nA x nB pair be held in the packed kernel's 12-bit score
// field, with the sentinel (-2000) below everything reachable and one extension short of wrapping?
// 0: no lower-case letter anywhere, 1: only in targets, 2: in queries (and maybe targets)
} // extern "C"
int igd_case_mix(const uint8_t *tgt, size_t tbytes, const uint8_t *qry, size_t qbytes)
{
    for (size_t i = 0; i < qbytes; ++i) if ((unsigned)(qry[i] - 'a') < 26u) return 2;
    for (size_t i = 0; i < tbytes; ++i) if ((unsigned)(tgt[i] - 'a') < 26u) return 1;
    return 0;
}
extern "C" {

} // extern "C"
bool igd_packed_scheme_ok(const igd_scoring *sc)
{
    const int g[] = {sc->target_internal_open, sc->target_internal_extend, sc->query_internal_open,
                     sc->query_internal_extend, sc->query_end_open, sc->query_end_extend};
    for (int v : g) if (v > 0 || v < -40) return false;
    return sc->match >= -40 && sc->match <= 40 && sc->mismatch >= -40 && sc->mismatch <= 40;
}
// 0: two-register kernel, 1: packed short layout, 2: packed long layout (see gotoh.cu)
int igd_packed_layout_for(const igd_scoring *sc, int nA, int nB)
{
    if (nB > 511) return 0;                                        // 9-bit matches field
    const int hi = std::max(0, sc->match) * std::min(nA, nB);
    // every state is at least as good as "row 0 along the query, then straight down" (pure gaps) ...
    int lo = sc->target_internal_open + sc->target_internal_extend * (nB - 1)
           + std::min(sc->query_internal_open, sc->query_end_open)
           + std::min(sc->query_internal_extend, sc->query_end_extend) * (nA - 1);
    // ... and with free query end gaps also as good as "free lead, then the diagonal" (or row 0 first)
    if (sc->query_end_open == 0 && sc->query_end_extend == 0)
        lo = std::max(lo, sc->target_internal_open + sc->target_internal_extend * (nB - 1)
                          + std::min(0, sc->mismatch) * std::min(nA, nB));
    // the short layout stores column j in a frame shifted by -j * te (gotoh.cu): values rise by up to nB * |te|
    if (nA <= 511 && hi + nB * std::max(0, -sc->target_internal_extend) <= 1900 && lo - 8 * 40 >= -1900) return 1;
    if (nA <= 1023 && hi <= 900 && lo - 2 * 40 >= -820) return 2;
    return 0;
}

extern "C" {

int igd_align_batch(igd_handle *h,
                    const uint8_t *tgt, const int32_t *tgt_off, int32_t n_tgt,
                    const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                    const int32_t *pair_t, const int32_t *pair_q, int64_t n_pairs,
                    const igd_scoring *sc,
                    int32_t *score, int32_t *matches, int32_t *alen, int32_t *start, int32_t *ient)
{
    if (!h || !tgt_off || !qry_off || !sc || n_tgt < 0 || n_qry < 0 || n_pairs < 0 || (n_pairs && (!pair_t || !pair_q || !score))) {
        igd_set_error("bad arguments to igd_align_batch"); return IGD_ERR_ARG;
    }
    if (n_pairs > 0x7fffffff) { igd_set_error("more than 2^31-1 pairs in one call"); return IGD_ERR_ARG; }
    if (sc->target_end_open != sc->target_internal_open || sc->target_end_extend != sc->target_internal_extend) {
        igd_set_error("target_end_* gap scores must equal target_internal_*"); return IGD_ERR_UNSUPPORTED;
    }
    if (sc->target_internal_extend < sc->target_internal_open || sc->query_internal_extend < sc->query_internal_open ||
        sc->query_end_extend < sc->query_end_open) {
        igd_set_error("gap extend scores must be >= gap open scores"); return IGD_ERR_UNSUPPORTED;
    }
    const int vals[] = {sc->match, sc->mismatch, sc->target_internal_open, sc->target_internal_extend,
                        sc->query_internal_open, sc->query_internal_extend, sc->query_end_open, sc->query_end_extend};
    for (int v : vals) if (v < -1000 || v > 1000) { igd_set_error("scores must lie in [-1000, 1000]"); return IGD_ERR_UNSUPPORTED; }
    IGD_CUDA(cudaSetDevice(h->device));
    h->timing.align_ms = 0.f; h->timing.align_launches = 0; h->timing.align_cells = 0;
    if (n_pairs == 0) return IGD_OK;

    // validate, bucket by target length and kernel, detect lower-case letters
    const size_t tbytes = (size_t)tgt_off[n_tgt], qbytes = (size_t)qry_off[n_qry];
    const int mixed = igd_case_mix(tgt, tbytes, qry, qbytes);
    const bool detail = start || ient;
    const bool packed_ok = !detail && igd_packed_scheme_ok(sc);
    constexpr int NB = 6;
    std::vector<int32_t> work[3][NB];                  // [kernel: 0 two-register, 1 packed short, 2 packed long][bucket]
    std::vector<int64_t> trivial;                      // empty target: closed form
    uint64_t cells = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        const int t = pair_t[p], q = pair_q[p];
        if (t < 0 || t >= n_tgt || q < 0 || q >= n_qry) { igd_set_error("pair %lld indexes outside the sequence tables", (long long)p); return IGD_ERR_ARG; }
        const int nA = tgt_off[t + 1] - tgt_off[t], nB = qry_off[q + 1] - qry_off[q];
        if (nA < 0 || nB < 1) { igd_set_error("pair %lld: query must be non-empty", (long long)p); return IGD_ERR_ARG; }
        if (nA > IGD_MAX_TARGET_LEN || nB > IGD_MAX_QUERY_LEN) {
            igd_set_error("pair %lld: lengths %d x %d exceed %d x %d", (long long)p, nA, nB, IGD_MAX_TARGET_LEN, IGD_MAX_QUERY_LEN);
            return IGD_ERR_TOO_LONG;
        }
        if (nA == 0) { trivial.push_back(p); continue; }
        const int b = igd_align_bucket_of(nA);
        int k = packed_ok ? igd_packed_layout_for(sc, nA, nB) : 0;
        if ((k == 1 && b > 3) || (k == 2 && b < 4)) k = 0;      // packed kernels exist for shapes 0-3 (short) and 4-5 (long)
        work[k][b].push_back((int32_t)p);
        cells += (uint64_t)nA * (uint64_t)nB;
    }
    int rc;
    if ((rc = igd_reserve(h, S_TGT, tbytes))) return rc;
    if ((rc = igd_reserve(h, S_TGTOFF, (size_t)(n_tgt + 1) * 4))) return rc;
    if ((rc = igd_reserve(h, S_QRY, qbytes))) return rc;
    if ((rc = igd_reserve(h, S_QRYOFF, (size_t)(n_qry + 1) * 4))) return rc;
    if ((rc = igd_reserve(h, S_PT, (size_t)n_pairs * 4))) return rc;
    if ((rc = igd_reserve(h, S_PQ, (size_t)n_pairs * 4))) return rc;
    if ((rc = igd_reserve(h, S_WORK, (size_t)n_pairs * 4))) return rc;
    if ((rc = igd_reserve(h, S_SCORE, (size_t)n_pairs * 4))) return rc;
    if ((rc = igd_reserve(h, S_PAY, (size_t)n_pairs * 4))) return rc;
    if (tbytes) IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TGT], tgt, tbytes, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TGTOFF], tgt_off, (size_t)(n_tgt + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRY], qry, qbytes, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRYOFF], qry_off, (size_t)(n_qry + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_PT], pair_t, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_PQ], pair_q, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, h->stream));
    // two-register kernel: flat lists of pair indices per bucket
    std::vector<int32_t> flat; flat.reserve((size_t)n_pairs);
    size_t woff[NB];
    for (int b = 0; b < NB; ++b) { woff[b] = flat.size(); flat.insert(flat.end(), work[0][b].begin(), work[0][b].end()); }
    if (!flat.empty())
        IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_WORK], flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, h->stream));
    // packed kernel: RUNS = one target against consecutive queries with consecutive output slots.
    // Maximal runs are cut into chunks small enough to balance the persistent warps (>= ~16 items each).
    std::vector<int4> runs[NB];                        // a bucket belongs to one packed layout (0-3 short, 4-5 long)
    size_t roff[NB], nruns_total = 0;
    for (int b = 0; b < NB; ++b) {
        const std::vector<int32_t> &w = work[b <= 3 ? 1 : 2][b];
        if (w.empty()) continue;
        const size_t slots = (size_t)h->sm_count * 16 * (b == 0 ? 4 : 1);       // concurrent groups of lanes
        const int chunk = (int)std::min<size_t>(64, std::max<size_t>(1, w.size() / (slots * 16)));
        for (size_t i = 0; i < w.size();) {
            const int32_t p = w[i];
            int n = 1;
            while (i + n < w.size() && n < chunk && w[i + n] == p + n && pair_t[p + n] == pair_t[p] && pair_q[p + n] == pair_q[p] + n) ++n;
            runs[b].push_back(make_int4(pair_t[p], pair_q[p], n, p));
            i += n;
        }
        nruns_total += runs[b].size();
    }
    if (nruns_total) {
        if ((rc = igd_reserve(h, S_RUNS, nruns_total * sizeof(int4)))) return rc;
        std::vector<int4> allruns; allruns.reserve(nruns_total);
        for (int b = 0; b < NB; ++b) { roff[b] = allruns.size(); allruns.insert(allruns.end(), runs[b].begin(), runs[b].end()); }
        IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_RUNS], allruns.data(), nruns_total * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
        IGD_CUDA(cudaStreamSynchronize(h->stream));          // allruns is a host temporary
    }

    std::vector<int32_t> h_score((size_t)n_pairs), h_pay((size_t)n_pairs);
    const int passes = ient ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
        IGD_CUDA(cudaMemsetAsync(h->d_count + 8, 0, 3 * NB * sizeof(unsigned long long), h->stream));
        IGD_CUDA(cudaEventRecord(h->ev0, h->stream));
        for (int k = 0; k < 3; ++k)
            for (int b = 0; b < NB; ++b) {
                if (work[k][b].empty()) continue;
                igd_align_job job;
                job.d_tgt = (const uint8_t *)h->d_scratch[S_TGT]; job.d_tgt_off = (const int32_t *)h->d_scratch[S_TGTOFF];
                job.d_qry = (const uint8_t *)h->d_scratch[S_QRY]; job.d_qry_off = (const int32_t *)h->d_scratch[S_QRYOFF];
                job.d_pair_t = (const int32_t *)h->d_scratch[S_PT]; job.d_pair_q = (const int32_t *)h->d_scratch[S_PQ];
                if (k == 0) {
                    job.d_work = (const int32_t *)h->d_scratch[S_WORK] + woff[b]; job.d_runs = nullptr;
                    job.n_work = (int32_t)work[0][b].size();
                } else {
                    job.d_work = nullptr; job.d_runs = (const int4 *)h->d_scratch[S_RUNS] + roff[b];
                    job.n_work = (int32_t)runs[b].size();
                }
                job.bucket = b; job.mixed_case = mixed; job.packed = k; job.pay_mode = pass; job.sc = *sc;
                job.d_score = (int32_t *)h->d_scratch[S_SCORE]; job.d_pay = (uint32_t *)h->d_scratch[S_PAY];
                job.d_counter = (unsigned int *)(h->d_count + 8 + k * NB + b);
                if ((rc = igd_launch_gotoh(h, job))) return rc;
            }
        IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
        IGD_CUDA(cudaMemcpyAsync(h_score.data(), h->d_scratch[S_SCORE], (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
        IGD_CUDA(cudaMemcpyAsync(h_pay.data(), h->d_scratch[S_PAY], (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
        IGD_CUDA(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        IGD_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->timing.align_ms += ms;
        h->timing.align_cells += cells;
        for (int64_t p = 0; p < n_pairs; ++p) {
            const int q = pair_q[p], t = pair_t[p];
            const int nA = tgt_off[t + 1] - tgt_off[t], nB = qry_off[q + 1] - qry_off[q];
            if (nA == 0) continue;
            const uint32_t pay = (uint32_t)h_pay[p];
            if (pass == 0) {
                score[p] = h_score[p];
                if (matches) matches[p] = (int)(pay & 1023);
                if (alen) alen[p] = nB + ((int)(pay >> 10) & 1023);
                if (start) start[p] = (int)(pay >> 20) & 1023;
            } else {
                ient[p] = nA - ((int)(pay >> 20) & 1023);
            }
        }
    }
    // empty target: the gene row is all letters against gaps (row 0 of the DP)
    for (int64_t p : trivial) {
        const int q = pair_q[p];
        const int nB = qry_off[q + 1] - qry_off[q];
        score[p] = sc->target_end_open + sc->target_end_extend * (nB - 1);
        if (matches) matches[p] = 0;
        if (alen) alen[p] = nB;
        if (start) start[p] = 0;
        if (ient) ient[p] = 0;
    }
    return IGD_OK;
}

// Dense form of igd_align_batch: every target against every query, outputs [n_tgt][n_qry].  This is
// the shape of align_fragment_to_genes (py/IGDetective.py:335-341) and of the iterative stage's
// ComputeAlignment (py/extract_aligned_genes.py:89-92); without pair tables the host does no per-pair
// work and every target becomes runs for the packed kernel.  Falls back to igd_align_batch when some
// pair does not fit the packed word.
} // extern "C"

// Every target against every query through the packed run-mode kernels, results left on the device
// (S_SCORE, S_PAY: [n_tgt][n_qry]); *launched = false (and nothing done) when some pair does not fit
// the packed kernels, the caller then takes the pair-table path.
int igd_dense_launch(igd_handle *h,
                     const uint8_t *tgt, const int32_t *tgt_off, int32_t n_tgt,
                     const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                     const igd_scoring *sc, bool *launched, bool tgt_resident)
{
    *launched = false;
    const int64_t n_pairs = (int64_t)n_tgt * (int64_t)n_qry;
    if (n_pairs > 0x7fffffff) { igd_set_error("more than 2^31-1 pairs in one call"); return IGD_ERR_ARG; }
    int max_q = 0, min_q = 1 << 30;
    for (int q = 0; q < n_qry; ++q) { const int n = qry_off[q + 1] - qry_off[q]; max_q = std::max(max_q, n); min_q = std::min(min_q, n); }
    bool fast = n_pairs > 0 && min_q >= 1 && max_q <= IGD_MAX_QUERY_LEN && igd_packed_scheme_ok(sc)
             && sc->target_end_open == sc->target_internal_open && sc->target_end_extend == sc->target_internal_extend
             && sc->target_internal_extend >= sc->target_internal_open && sc->query_internal_extend >= sc->query_internal_open
             && sc->query_end_extend >= sc->query_end_open;
    constexpr int NB = 6;
    std::vector<int32_t> tl[NB];                       // targets per kernel shape
    uint64_t cells = 0, qsum = n_qry ? (uint64_t)qry_off[n_qry] - (uint64_t)qry_off[0] : 0;
    for (int t = 0; t < n_tgt && fast; ++t) {
        const int nA = tgt_off[t + 1] - tgt_off[t];
        if (nA < 0 || nA > IGD_MAX_TARGET_LEN) { fast = false; break; }
        if (nA == 0) continue;
        const int b = igd_align_bucket_of(nA);
        // the packed word must hold the longest and the shortest query alike
        const int k = std::min(igd_packed_layout_for(sc, nA, max_q), igd_packed_layout_for(sc, nA, min_q)) == 0 ? 0
                    : std::max(igd_packed_layout_for(sc, nA, max_q), igd_packed_layout_for(sc, nA, min_q));
        if ((b <= 3 && k != 1) || (b >= 4 && k != 2)) { fast = false; break; }
        tl[b].push_back(t);
        cells += (uint64_t)nA * qsum;
    }
    if (!fast) return IGD_OK;
    IGD_CUDA(cudaSetDevice(h->device));
    h->timing.align_ms = 0.f; h->timing.align_launches = 0; h->timing.align_cells = 0;
    const size_t tbytes = (size_t)tgt_off[n_tgt], qbytes = (size_t)qry_off[n_qry];
    // resident targets (igd_detect: gathered and upper-cased on the device) hold no lower-case letter
    const int mixed = igd_case_mix(tgt, tgt_resident ? 0 : tbytes, qry, qbytes);
    int rc;
    if (!tgt_resident) {
        if ((rc = igd_reserve(h, S_TGT, tbytes))) return rc;
        if ((rc = igd_reserve(h, S_TGTOFF, (size_t)(n_tgt + 1) * 4))) return rc;
    }
    if ((rc = igd_reserve(h, S_QRY, qbytes))) return rc;
    if ((rc = igd_reserve(h, S_QRYOFF, (size_t)(n_qry + 1) * 4))) return rc;
    if ((rc = igd_reserve(h, S_SCORE, (size_t)n_pairs * 4))) return rc;
    if ((rc = igd_reserve(h, S_PAY, (size_t)n_pairs * 4))) return rc;
    if (tbytes && !tgt_resident) IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TGT], tgt, tbytes, cudaMemcpyHostToDevice, h->stream));
    if (!tgt_resident) IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_TGTOFF], tgt_off, (size_t)(n_tgt + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRY], qry, qbytes, cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_QRYOFF], qry_off, (size_t)(n_qry + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    // runs: each target's query list cut into chunks small enough to balance the persistent warps
    std::vector<int4> allruns;
    size_t roff[NB + 1];
    for (int b = 0; b < NB; ++b) {
        roff[b] = allruns.size();
        if (tl[b].empty()) continue;
        const size_t slots = (size_t)h->sm_count * 16 * (b == 0 ? 4 : 1);
        const size_t pairs_b = tl[b].size() * (size_t)n_qry;
        const int chunk = (int)std::min<size_t>(64, std::max<size_t>(1, pairs_b / (slots * 16)));
        for (int32_t t : tl[b])
            for (int q0 = 0; q0 < n_qry; q0 += chunk)
                allruns.push_back(make_int4(t, q0, std::min(chunk, n_qry - q0), (int)((size_t)t * n_qry + q0)));
    }
    roff[NB] = allruns.size();
    if ((rc = igd_reserve(h, S_RUNS, allruns.size() * sizeof(int4)))) return rc;
    IGD_CUDA(cudaMemcpyAsync(h->d_scratch[S_RUNS], allruns.data(), allruns.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    IGD_CUDA(cudaMemsetAsync(h->d_count + 8, 0, 3 * NB * sizeof(unsigned long long), h->stream));
    IGD_CUDA(cudaEventRecord(h->ev0, h->stream));
    for (int b = 0; b < NB; ++b) {
        if (tl[b].empty()) continue;
        igd_align_job job;
        job.d_tgt = (const uint8_t *)h->d_scratch[S_TGT]; job.d_tgt_off = (const int32_t *)h->d_scratch[S_TGTOFF];
        job.d_qry = (const uint8_t *)h->d_scratch[S_QRY]; job.d_qry_off = (const int32_t *)h->d_scratch[S_QRYOFF];
        job.d_pair_t = nullptr; job.d_pair_q = nullptr; job.d_work = nullptr;
        job.d_runs = (const int4 *)h->d_scratch[S_RUNS] + roff[b];
        job.n_work = (int32_t)(roff[b + 1] - roff[b]);
        job.bucket = b; job.mixed_case = mixed; job.packed = b <= 3 ? 1 : 2; job.pay_mode = 0; job.sc = *sc;
        job.d_score = (int32_t *)h->d_scratch[S_SCORE]; job.d_pay = (uint32_t *)h->d_scratch[S_PAY];
        job.d_counter = (unsigned int *)(h->d_count + 8 + job.packed * NB + b);
        if ((rc = igd_launch_gotoh(h, job))) return rc;
    }
    IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
    h->timing.align_cells = cells;
    *launched = true;
    return IGD_OK;
}

extern "C" {

int igd_align_dense(igd_handle *h,
                    const uint8_t *tgt, const int32_t *tgt_off, int32_t n_tgt,
                    const uint8_t *qry, const int32_t *qry_off, int32_t n_qry,
                    const igd_scoring *sc, int32_t *score, int32_t *matches, int32_t *alen)
{
    if (!h || !tgt_off || !qry_off || !sc || n_tgt < 0 || n_qry < 0 || !score) {
        igd_set_error("bad arguments to igd_align_dense"); return IGD_ERR_ARG;
    }
    const int64_t n_pairs = (int64_t)n_tgt * (int64_t)n_qry;
    if (n_pairs > 0x7fffffff) { igd_set_error("more than 2^31-1 pairs in one call"); return IGD_ERR_ARG; }
    bool launched = false;
    int rc;
    if ((rc = igd_dense_launch(h, tgt, tgt_off, n_tgt, qry, qry_off, n_qry, sc, &launched, false))) return rc;
    if (!launched) {                                   // generic path: explicit pair tables
        std::vector<int32_t> pt((size_t)n_pairs), pq((size_t)n_pairs);
        for (int t = 0; t < n_tgt; ++t)
            for (int q = 0; q < n_qry; ++q) { pt[(size_t)t * n_qry + q] = t; pq[(size_t)t * n_qry + q] = q; }
        return igd_align_batch(h, tgt, tgt_off, n_tgt, qry, qry_off, n_qry, pt.data(), pq.data(), n_pairs, sc,
                               score, matches, alen, nullptr, nullptr);
    }
    IGD_CUDA(cudaMemcpyAsync(score, h->d_scratch[S_SCORE], (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    std::vector<int32_t> h_pay((size_t)n_pairs);
    IGD_CUDA(cudaMemcpyAsync(h_pay.data(), h->d_scratch[S_PAY], (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    IGD_CUDA(cudaStreamSynchronize(h->stream));
    IGD_CUDA(cudaEventElapsedTime(&h->timing.align_ms, h->ev0, h->ev1));
    for (int t = 0; t < n_tgt; ++t) {
        const int nA = tgt_off[t + 1] - tgt_off[t];
        const size_t base = (size_t)t * n_qry;
        for (int q = 0; q < n_qry; ++q) {
            const int nB = qry_off[q + 1] - qry_off[q];
            if (nA == 0) {                              // gene row against gaps only (row 0 of the DP)
                score[base + q] = sc->target_end_open + sc->target_end_extend * (nB - 1);
                if (matches) matches[base + q] = 0;
                if (alen) alen[base + q] = nB;
                continue;
            }
            const uint32_t pay = (uint32_t)h_pay[base + q];
            if (matches) matches[base + q] = (int)(pay & 1023);
            if (alen) alen[base + q] = nB + ((int)(pay >> 10) & 1023);
        }
    }
    return IGD_OK;
}

} // extern "C"

// str(Seq(s).reverse_complement()): Bio.Seq's IUPAC table, case preserving (as fasta.py:_COMP);
// bytes outside the table stay as they are.
static uint8_t g_comp[256];
static std::once_flag g_comp_once;
static void init_comp()                                 // reachable from two handles on two threads
{
    std::call_once(g_comp_once, [] {
        for (int i = 0; i < 256; ++i) g_comp[i] = (uint8_t)i;
        const char *from = "ACGTMRWSYKVHDBXNacgtmrwsykvhdbxnUu", *to = "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxnAa";
        for (int i = 0; from[i]; ++i) g_comp[(uint8_t)from[i]] = (uint8_t)to[i];
    });
}

int igd_pin_reserve(igd_handle *h, int which, size_t bytes)
{
    if (h->h_pin_cap[which] >= bytes) return IGD_OK;
    if (h->h_pin[which]) { IGD_CUDA(cudaFreeHost(h->h_pin[which])); h->h_pin[which] = nullptr; h->h_pin_cap[which] = 0; }
    const size_t cap = bytes + bytes / 4 + 4096;
    IGD_CUDA(cudaMallocHost(&h->h_pin[which], cap));
    h->h_pin_cap[which] = cap;
    return IGD_OK;
}

int igd_stage_reserve(igd_handle *h, size_t bytes)
{
    if (h->h_stage_cap >= bytes) return IGD_OK;
    if (h->h_stage) { IGD_CUDA(cudaFreeHost(h->h_stage)); h->h_stage = nullptr; h->h_stage_cap = 0; }
    const size_t cap = bytes + bytes / 4 + 4096;
    IGD_CUDA(cudaMallocHost(&h->h_stage, cap));
    h->h_stage_cap = cap;
    return IGD_OK;
}

extern "C" {

int igd_score_fragments(igd_handle *h,
                        const uint8_t *frag, const int32_t *frag_off, int32_t n_frag,
                        const uint8_t *gene, const int32_t *gene_off, int32_t n_gene,
                        const igd_scoring *sc, double *pi, double *alen, uint8_t *direction)
{
    if (!h || !frag_off || !gene_off || !sc || n_frag < 0 || n_gene < 0 || !pi || !alen) {
        igd_set_error("bad arguments to igd_score_fragments"); return IGD_ERR_ARG;
    }
    if ((int64_t)n_frag * n_gene == 0) return IGD_OK;
    init_comp();
    // upper-cased fragments, identical ones scored once; targets 2u / 2u+1 = fragment u / its reverse complement
    std::unordered_map<std::string, int32_t> seen;
    std::vector<int32_t> uidx((size_t)n_frag, -1), toff(1, 0);
    std::vector<uint8_t> tgt;
    bool any_empty = false;
    for (int i = 0; i < n_frag; ++i) {
        const int n = frag_off[i + 1] - frag_off[i];
        if (n < 0) { igd_set_error("fragment offsets must ascend"); return IGD_ERR_ARG; }
        if (n == 0) { any_empty = true; continue; }
        std::string u((const char *)frag + frag_off[i], (size_t)n);
        for (char &c : u) if ((unsigned)(c - 'a') < 26u) c = (char)(c - 32);
        auto it = seen.find(u);
        if (it == seen.end()) {
            const int32_t id = (int32_t)seen.size();
            it = seen.emplace(std::move(u), id).first;
            const std::string &f = it->first;
            tgt.insert(tgt.end(), f.begin(), f.end());
            toff.push_back((int32_t)tgt.size());
            for (int k = n - 1; k >= 0; --k) tgt.push_back(g_comp[(uint8_t)f[(size_t)k]]);
            toff.push_back((int32_t)tgt.size());
        }
        uidx[(size_t)i] = it->second;
    }
    const int32_t U = (int32_t)seen.size();
    const size_t UG = (size_t)U * (size_t)n_gene;
    std::vector<uint32_t> host_pick;
    const uint32_t *pick = nullptr;
    float ms_total = 0.f; int launches = 0; uint64_t cells = 0;
    int rc;
    if (U > 0) {
        bool launched = false;
        // allocations first: nothing but kernel launches between the two timing events
        if ((rc = igd_reserve(h, S_PICK, UG * 4))) return rc;
        if ((rc = igd_stage_reserve(h, UG * 4))) return rc;
        if ((rc = igd_dense_launch(h, tgt.data(), toff.data(), 2 * U, gene, gene_off, n_gene, sc, &launched, false))) return rc;
        if (launched) {
            if ((rc = igd_launch_pick(h, (const uint32_t *)h->d_scratch[S_PAY], (const int32_t *)h->d_scratch[S_QRYOFF],
                                      U, n_gene, (uint32_t *)h->d_scratch[S_PICK]))) return rc;
            IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
            IGD_CUDA(cudaMemcpyAsync(h->h_stage, h->d_scratch[S_PICK], UG * 4, cudaMemcpyDeviceToHost, h->stream));
            IGD_CUDA(cudaStreamSynchronize(h->stream));
            IGD_CUDA(cudaEventElapsedTime(&h->timing.align_ms, h->ev0, h->ev1));
            pick = (const uint32_t *)h->h_stage;
        } else {                                       // some pair does not fit the packed kernels: pair-table path
            const size_t np = 2 * UG;
            std::vector<int32_t> s_(np), m_(np), l_(np);
            if ((rc = igd_align_dense(h, tgt.data(), toff.data(), 2 * U, gene, gene_off, n_gene, sc, s_.data(), m_.data(), l_.data()))) return rc;
            host_pick.resize(UG);
            for (int32_t u = 0; u < U; ++u)
                for (int32_t g = 0; g < n_gene; ++g) {
                    const size_t f = (size_t)(2 * u) * n_gene + g, r = (size_t)(2 * u + 1) * n_gene + g;
                    const bool fwd = m_[f] > m_[r];
                    const size_t w = fwd ? f : r;
                    host_pick[(size_t)u * n_gene + g] = (uint32_t)m_[w] | ((uint32_t)l_[w] << 10) | (fwd ? 0x80000000u : 0u);
                }
            pick = host_pick.data();
        }
        ms_total += h->timing.align_ms; launches += h->timing.align_launches; cells += h->timing.align_cells;
    }
    // an empty fragment stands for 'A' * len(gene) (py/IGDetective.py:339-340): one target pair per gene length
    std::vector<uint32_t> empty_pick;
    if (any_empty) {
        std::unordered_map<int, int32_t> by_len;
        std::vector<uint8_t> et; std::vector<int32_t> eoff(1, 0), pt, pq;
        for (int32_t g = 0; g < n_gene; ++g) {
            const int n = gene_off[g + 1] - gene_off[g];
            auto it = by_len.find(n);
            if (it == by_len.end()) {
                it = by_len.emplace(n, (int32_t)by_len.size()).first;
                et.insert(et.end(), (size_t)n, (uint8_t)'A'); eoff.push_back((int32_t)et.size());
                et.insert(et.end(), (size_t)n, (uint8_t)'T'); eoff.push_back((int32_t)et.size());
            }
            pt.push_back(2 * it->second); pq.push_back(g);
            pt.push_back(2 * it->second + 1); pq.push_back(g);
        }
        std::vector<int32_t> s_(pt.size()), m_(pt.size()), l_(pt.size());
        if ((rc = igd_align_batch(h, et.data(), eoff.data(), (int32_t)eoff.size() - 1, gene, gene_off, n_gene,
                                  pt.data(), pq.data(), (int64_t)pt.size(), sc, s_.data(), m_.data(), l_.data(), nullptr, nullptr))) return rc;
        empty_pick.resize((size_t)n_gene);
        for (int32_t g = 0; g < n_gene; ++g) {
            const bool fwd = m_[2 * (size_t)g] > m_[2 * (size_t)g + 1];
            const size_t w = 2 * (size_t)g + (fwd ? 0 : 1);
            empty_pick[(size_t)g] = (uint32_t)m_[w] | ((uint32_t)l_[w] << 10) | (fwd ? 0x80000000u : 0u);
        }
        ms_total += h->timing.align_ms; launches += h->timing.align_launches; cells += h->timing.align_cells;
    }
    h->timing.align_ms = ms_total; h->timing.align_launches = launches; h->timing.align_cells = cells;
    // BioAlign.PI(): (matches - 1) / len * 100 in float64, left to right (py/extract_aligned_genes.py:44-45)
    for (int i = 0; i < n_frag; ++i) {
        const uint32_t *row = uidx[(size_t)i] >= 0 ? pick + (size_t)uidx[(size_t)i] * n_gene : empty_pick.data();
        double *po = pi + (size_t)i * n_gene, *lo = alen + (size_t)i * n_gene;
        uint8_t *dd = direction ? direction + (size_t)i * n_gene : nullptr;
        for (int32_t g = 0; g < n_gene; ++g) {
            const uint32_t w = row[g];
            const double m = (double)((int)(w & 1023u) - 1), l = (double)((w >> 10) & 0x1fffffu);
            po[g] = m / l * 100.0;
            lo[g] = l;
            if (dd) dd[g] = (w & 0x80000000u) ? '+' : '-';
        }
    }
    return IGD_OK;
}

int igd_score_windows(igd_handle *h,
                      const uint8_t *win, const int32_t *win_off, int32_t n_win,
                      const uint8_t *gene, const int32_t *gene_off, int32_t n_gene,
                      const igd_scoring *sc, int32_t *best, double *pi, int32_t *start, int32_t *end, int32_t *ient)
{
    if (!h || !win_off || !gene_off || !sc || n_win < 0 || n_gene < 0 || !best || !pi || !start || !end || !ient) {
        igd_set_error("bad arguments to igd_score_windows"); return IGD_ERR_ARG;
    }
    for (int w = 0; w < n_win; ++w) { best[w] = -1; pi[w] = 0.0; start[w] = end[w] = ient[w] = 0; }
    if ((int64_t)n_win * n_gene == 0) return IGD_OK;
    init_comp();
    // targets 2w / 2w+1 = window w as it is (raw case) / its reverse complement
    const size_t wbytes = (size_t)win_off[n_win] - (size_t)win_off[0];
    std::vector<uint8_t> tgt(2 * wbytes);
    std::vector<int32_t> toff((size_t)2 * n_win + 1);
    size_t pos = 0;
    toff[0] = 0;
    for (int w = 0; w < n_win; ++w) {
        const int n = win_off[w + 1] - win_off[w];
        if (n < 0) { igd_set_error("window offsets must ascend"); return IGD_ERR_ARG; }
        const uint8_t *src = win + win_off[w];
        memcpy(tgt.data() + pos, src, (size_t)n); pos += (size_t)n;
        toff[(size_t)2 * w + 1] = (int32_t)pos;
        for (int k = n - 1; k >= 0; --k) tgt[pos++] = g_comp[src[k]];
        toff[(size_t)2 * w + 2] = (int32_t)pos;
    }
    std::vector<int4> sel((size_t)n_win);
    float ms_total = 0.f; int launches = 0; uint64_t cells = 0;
    bool launched = false;
    int rc;
    if ((rc = igd_reserve(h, S_PICK, (size_t)n_win * sizeof(int4)))) return rc;     // before the timing events
    if ((rc = igd_dense_launch(h, tgt.data(), toff.data(), 2 * n_win, gene, gene_off, n_gene, sc, &launched, false))) return rc;
    if (launched) {
        if ((rc = igd_launch_select_windows(h, (const uint32_t *)h->d_scratch[S_PAY], (const int32_t *)h->d_scratch[S_TGTOFF],
                                            (const int32_t *)h->d_scratch[S_QRYOFF], n_win, n_gene, (int4 *)h->d_scratch[S_PICK]))) return rc;
        IGD_CUDA(cudaEventRecord(h->ev1, h->stream));
        IGD_CUDA(cudaMemcpyAsync(sel.data(), h->d_scratch[S_PICK], (size_t)n_win * sizeof(int4), cudaMemcpyDeviceToHost, h->stream));
        IGD_CUDA(cudaStreamSynchronize(h->stream));
        IGD_CUDA(cudaEventElapsedTime(&h->timing.align_ms, h->ev0, h->ev1));
    } else {                                           // some pair does not fit the packed kernels: pair-table path
        const size_t np = (size_t)2 * n_win * n_gene;
        std::vector<int32_t> s_(np), m_(np), l_(np);
        if ((rc = igd_align_dense(h, tgt.data(), toff.data(), 2 * n_win, gene, gene_off, n_gene, sc, s_.data(), m_.data(), l_.data()))) return rc;
        for (int w = 0; w < n_win; ++w) {
            long long bn = 0, bd = 1; int bi = -1, bm = 0;
            for (int k = 0; k < 2 * n_gene; ++k) {
                const size_t i = (size_t)(2 * w + k / n_gene) * n_gene + (size_t)(k % n_gene);
                if (m_[i] < 1) continue;
                const long long num = m_[i] - 1, den = l_[i];
                if (bi < 0 || num * bd >= bn * den) { bn = num; bd = den; bi = k; bm = m_[i]; }
            }
            sel[(size_t)w] = make_int4(bi, bm, (int)bd, 0);
        }
    }
    ms_total += h->timing.align_ms; launches += h->timing.align_launches; cells += h->timing.align_cells;
    // the winner of each window is aligned once more for its range and query sequence (:95-99 use the BioAlign object)
    std::vector<int32_t> pt, pq, who;
    for (int w = 0; w < n_win; ++w) {
        const int k = sel[(size_t)w].x;
        if (k < 0) continue;
        pt.push_back(2 * w + k / n_gene); pq.push_back(k % n_gene); who.push_back(w);
    }
    if (!pt.empty()) {
        const size_t n = pt.size();
        std::vector<int32_t> s_(n), m_(n), l_(n), st_(n), ie_(n);
        if ((rc = igd_align_batch(h, tgt.data(), toff.data(), 2 * n_win, gene, gene_off, n_gene, pt.data(), pq.data(), (int64_t)n,
                                  sc, s_.data(), m_.data(), l_.data(), st_.data(), ie_.data()))) return rc;
        ms_total += h->timing.align_ms; launches += h->timing.align_launches; cells += h->timing.align_cells;
        for (size_t i = 0; i < n; ++i) {
            const int w = who[i];
            const int4 v = sel[(size_t)w];
            if (m_[i] != v.y || l_[i] != v.z) { igd_set_error("internal: packed and two-register kernels disagree"); return IGD_ERR_INTERNAL; }
            best[w] = v.x;
            pi[w] = (double)(v.y - 1) / (double)v.z * 100.0;
            start[w] = st_[i]; end[w] = st_[i] + l_[i]; ient[w] = ie_[i];
        }
    }
    h->timing.align_ms = ms_total; h->timing.align_launches = launches; h->timing.align_cells = cells;
    return IGD_OK;
}

} // extern "C"
