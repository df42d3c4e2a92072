// Mapping context and batch controller (C ABI of include/mapcaller_b200.h).
//
// Host side of the pipeline: owns the HBM-resident index replica, the arenas of one batch, the
// device-resident profile and the sequential state of the reference's thread body
// (src/ReadMapping.cpp:416-646): avgDist and the running totals.  The 200-read chunk protocol and the
// avgDist feedback (:538-539) are reproduced exactly by speculation: all chunks of a batch are mapped
// with the current EstiDistance, each pair reports the interval of EstiDistance values that leave its
// outcome unchanged, and the host walks the chunk statistics in file order, re-running only the chunks
// whose true EstiDistance falls outside their interval.
#include "mc_launch.h"
#ifndef MC_HOSTEMU
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is resolved at run time (see NcclApi)
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

void mc_set_error(const char* fmt, ...);

// ---- memory helpers ---------------------------------------------------------------------------------
#ifdef MC_HOSTEMU
#define MC_CHECK(x) do { } while (0)
static int dev_alloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : -1; }
static void dev_free(void* p) { free(p); }
static int host_alloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : -1; }
static void host_free(void* p) { free(p); }
static std::atomic<int64_t> g_h2d_bytes(0), g_d2h_bytes(0);
static int dev_h2d(void* d, const void* h, size_t n, mc_stream_t) { memcpy(d, h, n); return 0; }
static int dev_d2h(void* h, const void* d, size_t n, mc_stream_t) { memcpy(h, d, n); return 0; }
static int dev_d2d(void* d, const void* s, size_t n, mc_stream_t) { memcpy(d, s, n); return 0; }
static int dev_zero(void* d, size_t n, mc_stream_t) { memset(d, 0, n); return 0; }
static int dev_sync(mc_stream_t) { return 0; }
struct mc_event_t { int x; };
static void ev_create(mc_event_t*) {}
static void ev_destroy(mc_event_t*) {}
static void ev_record(mc_event_t*, mc_stream_t) {}
static double ev_ms(mc_event_t*, mc_event_t*) { return 0; }
#else
static int cuda_fail(cudaError_t e, const char* what)
{
	if (e == cudaSuccess) return 0;
	mc_set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
	return -1;
}
static int dev_alloc(void** p, size_t n) { return cuda_fail(cudaMalloc(p, n ? n : 1), "cudaMalloc"); }
static void dev_free(void* p) { if (p) cudaFree(p); }
static int host_alloc(void** p, size_t n) { return cuda_fail(cudaMallocHost(p, n ? n : 1), "cudaMallocHost"); }
static void host_free(void* p) { if (p) cudaFreeHost(p); }
static std::atomic<int64_t> g_h2d_bytes(0), g_d2h_bytes(0);   // what really crossed PCIe (reported by mc_get_stats)
static int dev_h2d(void* d, const void* h, size_t n, mc_stream_t s) { g_h2d_bytes += (int64_t)n; return n ? cuda_fail(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "H2D copy") : 0; }
static int dev_d2h(void* h, const void* d, size_t n, mc_stream_t s) { g_d2h_bytes += (int64_t)n; return n ? cuda_fail(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "D2H copy") : 0; }
static int dev_d2d(void* d, const void* sp, size_t n, mc_stream_t s) { return n ? cuda_fail(cudaMemcpyAsync(d, sp, n, cudaMemcpyDeviceToDevice, s), "D2D copy") : 0; }
static int dev_zero(void* d, size_t n, mc_stream_t s) { return n ? cuda_fail(cudaMemsetAsync(d, 0, n, s), "memset") : 0; }
static int dev_sync(mc_stream_t s)
{
	if (cuda_fail(cudaStreamSynchronize(s), "stream synchronize")) return -1;
	return cuda_fail(cudaGetLastError(), "kernel launch");
}
typedef cudaEvent_t mc_event_t;
static void ev_create(mc_event_t* e) { cudaEventCreate(e); }
static void ev_destroy(mc_event_t* e) { cudaEventDestroy(*e); }
static void ev_record(mc_event_t* e, mc_stream_t s) { cudaEventRecord(*e, s); }
static double ev_ms(mc_event_t* a, mc_event_t* b) { float ms = 0; cudaEventElapsedTime(&ms, *a, *b); return ms; }
#endif

#ifndef MC_HOSTEMU
// NCCL is bound with dlopen at the first multi-GPU call instead of at link time: a process that also imports torch must end
// up with ONE libnccl.so.2 (torch bundles a newer one than the system's and refuses to start on an older, already loaded copy).
struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	const char* (*GetErrorString)(ncclResult_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	bool ok;
};
static NcclApi* nccl_api()
{
	static NcclApi api; static bool tried = false;
	if (!tried)
	{
		tried = true; api.ok = false;
		void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (h)
		{
			*(void**)&api.GetUniqueId = dlsym(h, "ncclGetUniqueId"); *(void**)&api.CommInitRank = dlsym(h, "ncclCommInitRank");
			*(void**)&api.CommDestroy = dlsym(h, "ncclCommDestroy"); *(void**)&api.AllReduce = dlsym(h, "ncclAllReduce");
			*(void**)&api.AllGather = dlsym(h, "ncclAllGather"); *(void**)&api.Broadcast = dlsym(h, "ncclBroadcast");
			*(void**)&api.ReduceScatter = dlsym(h, "ncclReduceScatter");
			*(void**)&api.GetErrorString = dlsym(h, "ncclGetErrorString");
			*(void**)&api.GroupStart = dlsym(h, "ncclGroupStart"); *(void**)&api.GroupEnd = dlsym(h, "ncclGroupEnd");
			api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.ReduceScatter && api.Broadcast && api.GetErrorString && api.GroupStart && api.GroupEnd;
		}
	}
	if (!api.ok) { mc_set_error("libnccl.so.2 could not be loaded"); return nullptr; }
	return &api;
}
static bool nccl_api_loaded() { return nccl_api() != nullptr; }
static void nccl_destroy(ncclComm_t comm) { nccl_api()->CommDestroy(comm); }
#define ncclGetUniqueId nccl_api()->GetUniqueId
#define ncclCommInitRank nccl_api()->CommInitRank
#define ncclAllReduce nccl_api()->AllReduce
#define ncclAllGather nccl_api()->AllGather
#define ncclReduceScatter nccl_api()->ReduceScatter
#define ncclBroadcast nccl_api()->Broadcast
static int nccl_fail(ncclResult_t r, const char* what)
{
	if (r == ncclSuccess) return 0;
	mc_set_error("NCCL error in %s: %s", what, nccl_api()->GetErrorString(r));
	return -1;
}
#endif

struct DBuf {
	void* p = nullptr; size_t cap = 0;
	int reserve(size_t n) // contents are NOT preserved
	{
		if (n <= cap) return 0;
		dev_free(p); p = nullptr; cap = 0;
		size_t want = n + n / 4 + 256;
		if (dev_alloc(&p, want)) return -1;
		cap = want; return 0;
	}
	int grow_keep(size_t n, size_t used, mc_stream_t s)
	{
		if (n <= cap) return 0;
		void* q = nullptr; size_t want = n + n / 2 + 256;
		if (dev_alloc(&q, want)) return -1;
		if (used) { if (dev_d2d(q, p, used, s) || dev_sync(s)) return -1; }
		dev_free(p); p = q; cap = want; return 0;
	}
	void release() { dev_free(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return (T*)p; }
};
struct HBuf {
	void* p = nullptr; size_t cap = 0;
	int reserve(size_t n)
	{
		if (n <= cap) return 0;
		host_free(p); p = nullptr; cap = 0;
		size_t want = n + n / 4 + 256;
		if (host_alloc(&p, want)) return -1;
		cap = want; return 0;
	}
	void release() { host_free(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return (T*)p; }
};

enum { EV_START, EV_H2D, EV_SEED0, EV_SEED, EV_LOC0, EV_LOCATE, EV_CLUSTER, EV_PAIR0, EV_PAIR1, EV_ALN1, EV_PROF0, EV_PROF1, EV_D2H, EV_RED0, EV_RED1, EV_RST0, EV_RST1, EV_COUNT };

struct Bumps { mc_u64 pair, frag, aln, task, dpws, rtask, key, ptask, rwin; };
struct PersistBumps { mc_u64 bp, ind, ind_seq, pad; };

struct Staged { DBuf seq, roff, seed_off, cap, scan, flag; int64_t n_reads = 0, n_bytes = 0, n_slots = 0, base = 0; std::vector<int64_t> h_roff; bool valid = false;
	// a slot filled by mc_ingest_fastq keeps the FASTQ text and its newline table in HBM (mc_sam_text prints names, bases and qualities from them)
	DBuf text[2], lines[2]; int n_files = 0; bool has_text = false; int lpr = 4;
	const void* pre_src[2] = {nullptr, nullptr}; int64_t pre_len[2] = {0, 0}; bool prefetched = false;   // mc_ingest_prefetch
	int n_pieces = 0; int64_t piece_end[8];
	bool pending = false; };   // staged asynchronously: the consumer has to wait for ev_slot[] first   // read ranges [piece_end[p-1], piece_end[p]) whose bases arrive one after the other (events ev_piece[])

struct mc_ctx {
	mc_params prm;
	mc_stream_t stream;
	// index replica
	DBuf d_bwt, d_cbwt, d_sa, d_sa_dense, d_ktab, d_pac, d_chrom_end, d_chrom_id;
	DevIndex ix;
	int64_t G;
	// profile.  The arrays are padded to MC_TILE_PAD columns beyond G so that they divide into equal tiles for any number of ranks
	// (mc_profile_reduce_scatter).  own_beg / own_end: the columns whose counters are the library's - everything, until a
	// reduce-scatter leaves this rank with the sums of its tile only; own_carry = the difference-array sums of all columns before it.
	int64_t own_beg = 0, own_end = 0, own_tile = 0; int64_t own_carry[6] = {0, 0, 0, 0, 0, 0}; bool scattered = false;
	DBuf d_base16, d_sdiff, d_cdiff, d_mdiff, d_rcount, d_rflag;
	DBuf d_bp, d_ind, d_ind_seq, d_pbump;
	int64_t bp_cap = 0, ind_cap = 0, ind_seq_cap = 0;
	// batch arenas
	Staged cur; Staged slots[MC_SLOTS];
	DBuf d_slot_freq, d_seeds, d_slot_loc, d_loc_slot, d_pairs, d_npair;
	DBuf d_cands, d_ncand0, d_ncand, d_cscore, d_cpaired, d_corient, d_cfrag, d_cnfrag, d_ctmp;
	DBuf d_est, d_active, d_pair_flag, d_est_lo, d_est_hi, d_pair_out, d_chunk_out, d_chunk_lo, d_chunk_hi;
	DBuf d_rsum, d_frags, d_aln, d_tasks, d_dpws, d_rtask, d_rwin, d_rw_beg, d_bumps, d_stats, d_scan;
	DBuf d_rnp;
	DBuf d_keys, d_keys_tmp, d_accept, d_sort, d_read_redo, d_cap, d_ptask, d_disc, d_cand_off, d_scan2;
	HBuf h_disc;
	HBuf h_bounce[2], h_push; size_t push_at = 0;
	mc_stream_t cstream, dstream;   // copy + staging kernels; plain DMA of prefetched FASTQ blocks
#ifndef MC_HOSTEMU
	cudaEvent_t ev_bounce[2], ev_piece[8], ev_slot[MC_SLOTS], ev_text[MC_SLOTS];
#endif
	double frag_factor = 6.0, aln_factor = 3.0, dpws_factor = 2.0, task_factor = 1.0; int64_t rescue_cap = 1 << 20;
	// pinned host staging
	HBuf h_in_seq, h_in_off, h_seed_off, h_small, h_chunk, h_chunk_lo, h_chunk_hi, h_pairs, h_reads, h_cands, h_frags, h_aln, h_misc;
	std::vector<mc_read_out> reads_out; std::vector<mc_cand_out> cands_out;
	// results of mc_seed_cluster_batch
	std::vector<int64_t> sc_pair_off; std::vector<int32_t> sc_npairs, sc_nclusters; std::vector<SPair> sc_pairs; std::vector<Cand> sc_clusters;
	std::vector<mc_chunk_out> chunks_final;
	// sequential state
	mc_totals tot;
	bool discord_init = false; int64_t discord_gpos = 0, discord_dist = 0;
	PipeArgs last_a; int64_t last_frag_cap = 0; Bumps last_hb; bool last_profiled = true, defer_profile = false;   // the batch whose arenas are still on the device
	bool freeze_avg_dist = false;  // operator entry mc_rescue_batch: every chunk sees the avgDist the caller gave, nothing feeds back
	bool library_closed = false;   // a batch that was not a whole number of 200-read chunks has been mapped: only the last batch of a library may be
	std::vector<mc_site_rec> inv_sites, tnl_sites;
	// finalize products
	std::vector<mc_indel_rec> ind_out; std::vector<uint8_t> ind_seq_out; std::vector<mc_breakpoint_rec> bp_out;

	DBuf d_vc[18];   // scratch of mc_variant_scan, kept between calls
	HBuf h_vc_out, h_vc_depth; int64_t n_vc_out = 0;   // its result in page-locked memory
	DBuf d_ia[9]; HBuf h_ind_rec, h_ind_seq;   // scratch of mc_profile_indels
	int64_t last_n = 0; const int64_t* last_roff = nullptr;   // the batch whose arenas are still on the device (mc_sam_records)
	DBuf d_sam[5], d_chrom_names, d_chrom_name_off; HBuf h_sam_text; std::vector<mc_sam_rec> sam_out; std::vector<uint8_t> sam_cigar;
	std::vector<std::string> chrom_names;
	// stats
	mc_stats stats; DevStats dstats_last;
	mc_event_t ev[EV_COUNT];
#ifndef MC_HOSTEMU
	ncclComm_t comm = nullptr; int comm_rank = 0, comm_size = 1;
	DBuf d_comm_small, d_comm_buf, d_glist, d_glist_all; HBuf h_comm_buf;
#endif
	DBuf fq_cnt[2], fq_off[2], fq_scan, fq_rlen, fq_rsrc;   // scratch of mc_ingest_fastq
};

#define MC_TILE_ALIGN 25600          /* a multiple of MC_PROF_BLOCK (1024) and of MC_VC_BLOCK (100) */
#define MC_TILE_PAD (17 * MC_TILE_ALIGN) /* room for up to 16 equal tiles of whole MC_TILE_ALIGN units */
static void zero_stats(mc_stats* s) { memset(s, 0, sizeof(*s)); }

extern "C" {

void mc_params_default(mc_params* p)
{
	memset(p, 0, sizeof(*p));
	p->paired = 1; p->alg_ksw2 = 0; p->max_pos_diff = 30; p->max_clip = 5; p->max_dup = 5; p->max_mismatch_rate = 0.05f;
	p->update_profile = 1; p->want_alignments = 0; p->device = 0; p->shard_rank = 0; p->shard_count = 1;
}

void mc_ctx_destroy(mc_ctx* c)
{
#ifndef MC_HOSTEMU
	if (c) cudaSetDevice(c->prm.device);   // a process may hold contexts on several GPUs
#endif
	if (!c) return;
	DBuf* bufs[] = {&c->d_bwt, &c->d_cbwt, &c->d_sa, &c->d_sa_dense, &c->d_ktab, &c->d_pac, &c->d_chrom_end, &c->d_chrom_id, &c->d_base16, &c->d_sdiff, &c->d_cdiff, &c->d_mdiff, &c->d_rcount, &c->d_rflag, &c->d_bp, &c->d_ind,
	                &c->d_ind_seq, &c->d_pbump, &c->d_slot_freq, &c->d_seeds, &c->d_slot_loc, &c->d_loc_slot, &c->d_pairs, &c->d_npair, &c->d_cands,
	                &c->d_ncand0, &c->d_ncand, &c->d_cscore, &c->d_cpaired, &c->d_corient, &c->d_cfrag, &c->d_cnfrag, &c->d_ctmp, &c->d_est, &c->d_active,
	                &c->d_pair_flag, &c->d_est_lo, &c->d_est_hi, &c->d_pair_out, &c->d_chunk_out, &c->d_chunk_lo, &c->d_chunk_hi, &c->d_rsum, &c->d_frags,
	                &c->d_aln, &c->d_tasks, &c->d_dpws, &c->d_rtask, &c->d_rwin, &c->d_rw_beg, &c->d_bumps, &c->d_stats, &c->d_scan, &c->d_keys, &c->d_keys_tmp, &c->d_accept, &c->d_rnp, &c->d_sort, &c->d_read_redo, &c->d_cap, &c->d_ptask, &c->d_disc, &c->d_cand_off, &c->d_scan2};
	for (DBuf* b : bufs) b->release();
	for (DBuf& b : c->d_vc) b.release();
	for (DBuf& b : c->d_ia) b.release();
	c->h_ind_rec.release(); c->h_ind_seq.release(); c->h_vc_out.release(); c->h_vc_depth.release();
	for (DBuf& b : c->d_sam) b.release();
	c->d_chrom_names.release(); c->d_chrom_name_off.release(); c->h_sam_text.release();
	std::vector<Staged*> st; st.push_back(&c->cur); for (int i = 0; i < MC_SLOTS; i++) st.push_back(&c->slots[i]);
	for (Staged* s : st) { s->seq.release(); s->roff.release(); s->seed_off.release(); s->cap.release(); s->scan.release(); s->flag.release(); for (int f = 0; f < 2; f++) { s->text[f].release(); s->lines[f].release(); } }
	DBuf* fq[] = {&c->fq_cnt[0], &c->fq_cnt[1], &c->fq_off[0], &c->fq_off[1], &c->fq_scan, &c->fq_rlen, &c->fq_rsrc};
	for (DBuf* b : fq) b->release();
	HBuf* hb[] = {&c->h_in_seq, &c->h_in_off, &c->h_seed_off, &c->h_small, &c->h_chunk, &c->h_chunk_lo, &c->h_chunk_hi, &c->h_pairs, &c->h_reads, &c->h_cands, &c->h_frags, &c->h_aln, &c->h_misc, &c->h_disc};
	for (HBuf* b : hb) b->release();
#ifndef MC_HOSTEMU
	if (c->comm && nccl_api_loaded()) nccl_destroy(c->comm);
	c->d_comm_small.release(); c->d_comm_buf.release(); c->d_glist.release(); c->d_glist_all.release(); c->h_comm_buf.release();
#endif
	for (int i = 0; i < EV_COUNT; i++) ev_destroy(&c->ev[i]);
	c->h_bounce[0].release(); c->h_bounce[1].release(); c->h_push.release();
#ifndef MC_HOSTEMU
	cudaEventDestroy(c->ev_bounce[0]); cudaEventDestroy(c->ev_bounce[1]); for (int i = 0; i < 8; i++) cudaEventDestroy(c->ev_piece[i]);
	for (int i = 0; i < MC_SLOTS; i++) { cudaEventDestroy(c->ev_slot[i]); cudaEventDestroy(c->ev_text[i]); }
	if (c->dstream) cudaStreamDestroy(c->dstream);
	if (c->cstream) cudaStreamDestroy(c->cstream);
	if (c->stream) cudaStreamDestroy(c->stream);
#endif
	delete c;
}

int mc_ctx_create(const mc_index* idx, const mc_params* params, mc_ctx** out)
{
	if (!idx || !params || !out) { mc_set_error("mc_ctx_create: null argument"); return MC_ERR_ARG; }
	if (params->max_dup < 1 || params->max_dup > 15 || params->max_pos_diff < 0) { mc_set_error("mc_ctx_create: parameter out of range"); return MC_ERR_ARG; }
	mc_index_view v;
	if (mc_index_get(idx, &v)) return MC_ERR_ARG;
	mc_ctx* c = new mc_ctx();
	c->prm = *params; memset(&c->tot, 0, sizeof(c->tot)); c->tot.avg_dist = 1000; zero_stats(&c->stats);
	for (int i = 0; i < v.n_chrom; i++) c->chrom_names.push_back(v.chrom_name && v.chrom_name[i] ? v.chrom_name[i] : "*");
#ifndef MC_HOSTEMU
	int ndev = 0;
	if (cuda_fail(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev <= params->device)
	{
		if (ndev <= params->device && ndev > 0) mc_set_error("mc_ctx_create: CUDA device %d not present (%d visible)", params->device, ndev);
		delete c; return MC_ERR_CUDA;
	}
	if (cuda_fail(cudaSetDevice(params->device), "cudaSetDevice")) { delete c; return MC_ERR_CUDA; }
	// the mapping stream outranks the staging streams: a batch being parsed for later must not hold up the kernels of the batch
	// being mapped now
	int prio_lo = 0, prio_hi = 0; cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
	if (cuda_fail(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate")) { delete c; return MC_ERR_CUDA; }
	if (cuda_fail(cudaStreamCreateWithPriority(&c->cstream, cudaStreamNonBlocking, prio_lo), "cudaStreamCreate")) { delete c; return MC_ERR_CUDA; }
	if (cuda_fail(cudaStreamCreateWithPriority(&c->dstream, cudaStreamNonBlocking, prio_lo), "cudaStreamCreate")) { delete c; return MC_ERR_CUDA; }
#else
	c->stream = 0; c->cstream = 0; c->dstream = 0;
#endif
	for (int i = 0; i < EV_COUNT; i++) ev_create(&c->ev[i]);
#ifndef MC_HOSTEMU
	cudaEventCreateWithFlags(&c->ev_bounce[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_bounce[1], cudaEventDisableTiming);
	for (int i = 0; i < 8; i++) cudaEventCreateWithFlags(&c->ev_piece[i], cudaEventDisableTiming);
	for (int i = 0; i < MC_SLOTS; i++) { cudaEventCreateWithFlags(&c->ev_slot[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_text[i], cudaEventDisableTiming); }
#endif
	const int64_t G = v.genome_size; c->G = G;
	std::vector<int64_t> ends; std::vector<int32_t> ids;
	{
		// PosChrIdMap (reference src/bwt_index.cpp:244-255): forward ends ascending, then reverse-strand ends
		std::vector<std::pair<int64_t, int32_t> > kv; int64_t tot = 0;
		for (int i = 0; i < v.n_chrom; i++)
		{
			int64_t fwd = tot; tot += v.chrom_len[i]; int64_t rev = 2 * G - tot;
			kv.push_back(std::make_pair(fwd + v.chrom_len[i] - 1, i)); kv.push_back(std::make_pair(rev + v.chrom_len[i] - 1, i));
		}
		std::sort(kv.begin(), kv.end());
		for (size_t i = 0; i < kv.size(); i++) { ends.push_back(kv[i].first); ids.push_back(kv[i].second); }
	}
	const size_t pac_bytes = (size_t)(G / 4 + 1);
	int bad = 0;
	bad |= c->d_bwt.reserve(v.bwt_size * 4 + 64) || dev_h2d(c->d_bwt.p, v.bwt, v.bwt_size * 4, c->stream);
	bad |= c->d_sa.reserve(v.n_sa * 8) || dev_h2d(c->d_sa.p, v.sa, v.n_sa * 8, c->stream);
	bad |= c->d_pac.reserve(pac_bytes + 16) || dev_h2d(c->d_pac.p, v.pac, pac_bytes, c->stream);
	bad |= c->d_chrom_end.reserve(ends.size() * 8) || dev_h2d(c->d_chrom_end.p, ends.data(), ends.size() * 8, c->stream);
	bad |= c->d_chrom_id.reserve(ids.size() * 4) || dev_h2d(c->d_chrom_id.p, ids.data(), ids.size() * 4, c->stream);
	if (params->update_profile)
	{
		const size_t GP = (size_t)G + MC_TILE_PAD;
		bad |= c->d_base16.reserve(GP * 8) || c->d_sdiff.reserve(GP * 16) || c->d_cdiff.reserve(GP * 4);
		bad |= c->d_mdiff.reserve(GP * 4) || c->d_rcount.reserve(GP);
		bad |= dev_zero(c->d_base16.p, GP * 8, c->stream) || dev_zero(c->d_sdiff.p, GP * 16, c->stream) || dev_zero(c->d_cdiff.p, GP * 4, c->stream);
		bad |= dev_zero(c->d_mdiff.p, GP * 4, c->stream) || dev_zero(c->d_rcount.p, GP, c->stream);
		c->own_beg = 0; c->own_end = G;
	}
	bad |= c->d_pbump.reserve(sizeof(PersistBumps)) || dev_zero(c->d_pbump.p, sizeof(PersistBumps), c->stream);
	bad |= c->d_bumps.reserve(sizeof(Bumps)) || c->d_stats.reserve(sizeof(DevStats)) || dev_zero(c->d_stats.p, sizeof(DevStats), c->stream);
	bad |= c->h_small.reserve(4096);
	bad |= dev_sync(c->stream);
	if (bad) { mc_ctx_destroy(c); return MC_ERR_CUDA; }
	DevIndex& ix = c->ix;
	ix.cbwt = nullptr;
	if (v.seq_len < (1ull << 32) && !params->reserved[1])
	{
		// re-block the index into the compact 32-byte layout (mc_fmindex.h); the reference-layout copy is released afterwards
		const int64_t ncb = (int64_t)((v.seq_len + 63) / 64) + 1;
		if (c->d_cbwt.reserve((size_t)ncb * 32 + 64) || dev_zero(c->d_cbwt.p, (size_t)ncb * 32 + 64, c->stream)) { mc_ctx_destroy(c); return MC_ERR_CUDA; }
		launch_cbwt_build(ncb - 1, c->d_bwt.as<uint32_t>(), c->d_cbwt.as<uint32_t>(), c->stream);
		if (dev_sync(c->stream)) { mc_ctx_destroy(c); return MC_ERR_CUDA; }
		ix.cbwt = c->d_cbwt.as<uint32_t>();
		c->d_bwt.release();
	}
	ix.bwt = c->d_bwt.as<uint32_t>(); ix.sa = c->d_sa.as<uint64_t>(); ix.pac = c->d_pac.as<uint8_t>();
	ix.chrom_end = c->d_chrom_end.as<int64_t>(); ix.chrom_id = c->d_chrom_id.as<int32_t>(); ix.n_end = (int32_t)ends.size();
	ix.primary = v.primary; for (int i = 0; i < 5; i++) ix.L2[i] = v.L2[i]; ix.seq_len = v.seq_len; ix.G = G; ix.twoG = 2 * G;
	ix.sa32 = nullptr; ix.sa_shift = 5; ix.ktab32 = nullptr; ix.ktab64 = nullptr; ix.ktab_k = 0;
	{
		// k-mer start table of the seed search (mc_fmindex.h): k = floor(log4(text length)) - 2, i.e. k-mers that still occur a few
		// dozen times on average, so that nearly every lookup succeeds.  MC_KMER_K overrides (0 = no table).
		int k = 0; for (uint64_t t = v.seq_len; t >= 4; t >>= 2) k++;
		k -= 2; if (k > 14) k = 14; if (k < 4) k = 4;
		if (getenv("MC_KMER_K")) k = atoi(getenv("MC_KMER_K"));
		if (k >= 2 && k <= 15)
		{
			const bool narrow = ix.cbwt != nullptr;
			const size_t ne = (size_t)1 << (2 * k);
			if (c->d_ktab.reserve(ne * (narrow ? 8 : 16))) { mc_ctx_destroy(c); return MC_ERR_CUDA; }
			launch_ktab_build(ix, k, narrow ? c->d_ktab.as<uint32_t>() : nullptr, narrow ? nullptr : c->d_ktab.as<uint64_t>(), c->stream);
			if (dev_sync(c->stream)) { mc_ctx_destroy(c); return MC_ERR_CUDA; }
			if (narrow) ix.ktab32 = c->d_ktab.as<uint32_t>(); else ix.ktab64 = c->d_ktab.as<uint64_t>();
			ix.ktab_k = k;
		}
	}
	{
		// denser sample of the suffix array in HBM (mc_fmindex.h): every 4th row, 32-bit, for texts below 2^32 symbols; every 8th
		// row, 64-bit, otherwise.  MC_SA_SHIFT=5 keeps the reference's every-32nd-row sample (A/B measurements).
		const bool narrow = v.seq_len < (1ull << 32) && !params->reserved[1];   // reserved[1]: the whole >= 2^32 path, also on a small text (tests)
		int shift = narrow ? 2 : 3;
		if (getenv("MC_SA_SHIFT")) shift = atoi(getenv("MC_SA_SHIFT"));
		if (shift >= 0 && shift < 5)
		{
			const int64_t nd = (int64_t)(v.seq_len >> shift) + 1;
			if (c->d_sa_dense.reserve((size_t)nd * (narrow ? 4 : 8))) { mc_ctx_destroy(c); return MC_ERR_CUDA; }
			launch_sa_dense(ix, nd, shift, narrow ? c->d_sa_dense.as<uint32_t>() : nullptr, narrow ? nullptr : c->d_sa_dense.as<uint64_t>(), c->stream);
			if (dev_sync(c->stream)) { mc_ctx_destroy(c); return MC_ERR_CUDA; }
			if (narrow) ix.sa32 = c->d_sa_dense.as<uint32_t>(); else ix.sa = c->d_sa_dense.as<uint64_t>();
			ix.sa_shift = shift;
			c->d_sa.release();
		}
	}
	*out = c;
	return MC_OK;
}

int mc_reset(mc_ctx* c)
{
#ifndef MC_HOSTEMU
	if (c) cudaSetDevice(c->prm.device);   // a process may hold contexts on several GPUs
#endif
	if (!c) { mc_set_error("mc_reset: null context"); return MC_ERR_ARG; }
	int bad = 0;
	ev_record(&c->ev[EV_RST0], c->stream);
	if (c->prm.update_profile)
	{
		const size_t GP = (size_t)c->G + (c->scattered ? MC_TILE_PAD : 1);   // the padding only ever holds something after a reduce-scatter
		bad |= dev_zero(c->d_base16.p, GP * 8, c->stream) || dev_zero(c->d_sdiff.p, GP * 16, c->stream) || dev_zero(c->d_cdiff.p, GP * 4, c->stream);
		bad |= dev_zero(c->d_mdiff.p, GP * 4, c->stream) || dev_zero(c->d_rcount.p, GP, c->stream);
		c->own_beg = 0; c->own_end = c->G; c->scattered = false; for (int k = 0; k < 6; k++) c->own_carry[k] = 0;
	}
	bad |= dev_zero(c->d_pbump.p, sizeof(PersistBumps), c->stream);
	ev_record(&c->ev[EV_RST1], c->stream);
	bad |= dev_sync(c->stream);
	if (!bad) { const double ms = ev_ms(&c->ev[EV_RST0], &c->ev[EV_RST1]); c->stats.ms_reset += ms; c->stats.ms_total += ms; }
	memset(&c->tot, 0, sizeof(c->tot)); c->tot.avg_dist = 1000;
	c->inv_sites.clear(); c->tnl_sites.clear(); c->discord_gpos = c->discord_dist = 0; c->library_closed = false;
	return bad ? MC_ERR_CUDA : MC_OK;
}

// The loop over libraries of Mapping() (reference src/ReadMapping.cpp:705-748): every library is cut into 200-read chunks from
// its own start; the profile, the totals and avgDist carry over.
int mc_begin_library(mc_ctx* c)
{
	if (!c) { mc_set_error("mc_begin_library: null context"); return MC_ERR_ARG; }
	c->library_closed = false;
	return MC_OK;
}

int mc_host_alloc(size_t bytes, void** out)
{
	if (!out) { mc_set_error("mc_host_alloc: null argument"); return MC_ERR_ARG; }
	return host_alloc(out, bytes) ? MC_ERR_CUDA : MC_OK;
}
void mc_host_free(void* p) { host_free(p); }

int mc_get_totals(const mc_ctx* c, mc_totals* out) { if (!c || !out) return MC_ERR_ARG; *out = c->tot; return MC_OK; }
int mc_set_totals(mc_ctx* c, const mc_totals* in) { if (!c || !in) return MC_ERR_ARG; c->tot = *in; return MC_OK; }
int mc_get_stats(const mc_ctx* c, mc_stats* out) { if (!c || !out) return MC_ERR_ARG; *out = c->stats; out->kernel_launches = g_launches; out->h2d_bytes = g_h2d_bytes; out->d2h_bytes = g_d2h_bytes; return MC_OK; }
int mc_reset_stats(mc_ctx* c) { if (!c) return MC_ERR_ARG; zero_stats(&c->stats); g_launches = 0; g_h2d_bytes = 0; g_d2h_bytes = 0; return MC_OK; }

} // extern "C"

// ---- staging ------------------------------------------------------------------------------------------
// Host -> device copy of caller memory.  Pinned / registered memory (e.g. from mc_host_alloc) is handed to the copy engine
// directly; pageable memory is streamed through two pinned bounce buffers so that the host memcpy of one piece overlaps the
// DMA of the previous one.
static int upload(mc_ctx* c, void* dst, const void* src, size_t bytes, mc_stream_t stream)
{
	if (!bytes) return 0;
#ifndef MC_HOSTEMU
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, src) == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged))
	{
		// in pieces of 4 MB: the copy engine takes its commands in order, and a batch being mapped sends a few kilobytes host ->
		// device per attempt - behind one 600 MB transfer those would wait 10+ ms each (measured: mapping ran at 40 % of its speed
		// while a prefetch was on the wire), behind a 4 MB piece they wait 0.1 ms
		const size_t piece = (size_t)4 << 20;
		for (size_t off = 0; off < bytes; off += piece)
			if (dev_h2d((uint8_t*)dst + off, (const uint8_t*)src + off, std::min(piece, bytes - off), stream)) return -1;
		return 0;
	}
	cudaGetLastError();
	const size_t piece = 8u << 20;
	if (c->h_bounce[0].reserve(piece) || c->h_bounce[1].reserve(piece)) return -1;
	for (size_t off = 0, k = 0; off < bytes; off += piece, k++)
	{
		const size_t m = std::min(piece, bytes - off);
		const int slot = (int)(k & 1);
		// the slot may still be feeding an earlier copy (also one issued by a previous upload() call)
		if (cuda_fail(cudaEventSynchronize(c->ev_bounce[slot]), "bounce buffer wait")) return -1;
		memcpy(c->h_bounce[slot].p, (const uint8_t*)src + off, m);
		if (dev_h2d((uint8_t*)dst + off, c->h_bounce[slot].p, m, stream)) return -1;
		cudaEventRecord(c->ev_bounce[slot], stream);
	}
	return 0;
#else
	return dev_h2d(dst, src, bytes, stream);
#endif
}

// Small host -> device transfers of the batch controller (a few kilobytes per attempt).  They do not go through the copy
// engine: that engine takes its commands in submission order, so behind the FASTQ blocks mc_ingest_prefetch queued for the next
// batches (1.3 GB each) they waited for tens of milliseconds and the mapping ran at 40 % of its speed.  The bytes are placed in
// a page-locked ring and a one-block kernel reads them over PCIe itself.
#ifndef MC_HOSTEMU
__global__ void mc_push_kernel(uint8_t* dst, const uint8_t* src, size_t n)
{ for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i]; }
#endif
static int dev_push(mc_ctx* c, void* dst, const void* src, size_t bytes, mc_stream_t s)
{
	if (!bytes) return 0;
#ifdef MC_HOSTEMU
	memcpy(dst, src, bytes); return 0;
#else
	const size_t ring = (size_t)8 << 20;
	if (bytes > ring / 4) return dev_h2d(dst, src, bytes, s);
	if (c->h_push.reserve(ring)) return -1;
	size_t at = (c->push_at + 255) & ~(size_t)255;
	if (at + bytes > ring) at = 0;      // a slot is reused 8 MB of pushes later: many batches, each of which ended with a synchronize
	c->push_at = at + bytes;
	memcpy(c->h_push.as<uint8_t>() + at, src, bytes);
	g_h2d_bytes += (int64_t)bytes;
	mc_push_kernel<<<(unsigned)std::min<size_t>((bytes + 255) / 256, 64), 256, 0, s>>>((uint8_t*)dst, c->h_push.as<uint8_t>() + at, bytes);
	return cuda_fail(cudaGetLastError(), "push kernel");
#endif
}

struct CapArgs { const int64_t* roff; uint32_t* cap; mc_u64* flag; };   // flag: the slot's own copy of DevStats::overflow (staging may run beside a batch)
MC_HD void seedcap_body(int64_t r, const CapArgs& q)
{
	const int64_t len = q.roff[r + 1] - q.roff[r];
	if (len < 0 || len > MC_MAX_RLEN) { mc_atomic_or(q.flag, (mc_u64)1 << 56); q.cap[r] = 1; return; }
	q.cap[r] = (uint32_t)(len / 17 + 1);   // a recorded seed is >= 16 bases and the next search starts one base later
}
#ifdef MC_HOSTEMU
static void launch_seedcap(const CapArgs& q, int64_t n, mc_stream_t) { for (int64_t i = 0; i < n; i++) seedcap_body(i, q); }
#else
__global__ void __launch_bounds__(MC_BLOCK) mc_seedcap_kernel(const CapArgs q, int64_t n)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) seedcap_body(i, q); }
static void launch_seedcap(const CapArgs& q, int64_t n, mc_stream_t s)
{ if (n > 0) { mc_seedcap_kernel<<<(unsigned)((n + MC_BLOCK - 1) / MC_BLOCK), MC_BLOCK, 0, s>>>(q, n); g_launches++; } }
#endif

// `pieces` > 1: the bases travel in that many chunk-aligned pieces, each followed by an event, so that the consumer can start
// on a piece while the rest is still on the wire (the offsets and everything derived from them go first)
static int stage_reads(mc_ctx* c, const mc_batch_in* in, Staged& st, mc_stream_t stream, int pieces)
{
	const int64_t n = in->n_reads;
	if (n < 0 || (n > 0 && (!in->seq || !in->seq_off))) { mc_set_error("mc_map_batch: bad batch"); return MC_ERR_ARG; }
	if (c->prm.paired && (n & 1)) { mc_set_error("mc_map_batch: paired mode needs an even number of reads"); return MC_ERR_ARG; }
	if (n >= (1ll << MC_KEY_SHIFT)) { mc_set_error("mc_map_batch: at most %lld reads per batch", (1ll << MC_KEY_SHIFT) - 1); return MC_ERR_ARG; }
	const int64_t base = n ? in->seq_off[0] : 0, bytes = n ? in->seq_off[n] - base : 0;
	if (bytes < 0 || bytes > n * (int64_t)MC_MAX_RLEN) { mc_set_error("mc_map_batch: read offsets are not increasing or a read exceeds %d bases", MC_MAX_RLEN); return MC_ERR_ARG; }
	st.n_reads = n; st.n_bytes = bytes; st.base = base; st.n_slots = bytes / 17 + n;   // upper bound of the seed slots
	st.n_pieces = 0;
	st.h_roff.clear();
	if (c->prm.want_alignments) st.h_roff.assign(in->seq_off, in->seq_off + n + 1);
	if (n == 0) { st.valid = true; return MC_OK; }
	// scratch of the staging pass is private to the Staged slot
	if (st.seq.reserve(bytes + 16) || st.roff.reserve((n + 1) * 8) || st.seed_off.reserve((n + 2) * 8) || st.cap.reserve(n * 4) || st.scan.reserve(device_scan_scratch_bytes(n)) || st.flag.reserve(16)) return MC_ERR_CUDA;
	if (dev_zero(st.flag.p, 16, stream) || upload(c, st.roff.p, in->seq_off, (n + 1) * 8, stream)) return MC_ERR_CUDA;
	CapArgs q; q.roff = st.roff.as<int64_t>(); q.cap = st.cap.as<uint32_t>(); q.flag = st.flag.as<mc_u64>();
	launch_seedcap(q, n, stream);
	device_scan_u32(q.cap, st.seed_off.as<int64_t>(), n, st.scan.as<int64_t>(), stream);
#ifndef MC_HOSTEMU
	if (pieces > 1)
	{
		const int64_t n_chunks = (n + MC_CHUNK_READS - 1) / MC_CHUNK_READS;
		const int64_t per = ((n_chunks + pieces - 1) / pieces) * MC_CHUNK_READS;
		for (int p = 0; p < pieces; p++)
		{
			const int64_t r0 = std::min(n, (int64_t)p * per), r1 = std::min(n, (int64_t)(p + 1) * per);
			if (upload(c, st.seq.as<uint8_t>() + (in->seq_off[r0] - base), in->seq + in->seq_off[r0], (size_t)(in->seq_off[r1] - in->seq_off[r0]), stream)) return MC_ERR_CUDA;
			if (cuda_fail(cudaEventRecord(c->ev_piece[p], stream), "cudaEventRecord")) return MC_ERR_CUDA;
			st.piece_end[p] = r1;
		}
		st.n_pieces = pieces;
		st.valid = true;
		return MC_OK;
	}
#endif
	if (upload(c, st.seq.p, in->seq + base, bytes, stream)) return MC_ERR_CUDA;
	st.valid = true;
	return MC_OK;
}

// ---- ordered exchange between the GPUs of one library ------------------------------------------------------
// After mc_comm_init() the ranks map consecutive shards of ONE library: per mc_map_batch call (a collective), rank r's reads
// follow those of rank r-1 in file order, and the next call continues after the last rank.  What the reference does
// sequentially in file order is exchanged so that the result equals one reference thread over the whole library
// (SURVEY.md section 8e):
//   * avgDist feedback: the per-chunk sums and validity intervals of every rank are all-gathered after each attempt and
//     every rank walks the global chunk sequence (same decisions everywhere, each rank re-runs its own chunks);
//   * dedup gate: per-start candidate counts of the ranks before / after this one are added to readCount around the gate;
//   * discordant pairs: the few records are gathered and replayed in global order, every rank keeps its own sites.
// mc_params.reserved[2] = 1 turns the exchange off (independent shards, results summed by mc_profile_allreduce).
struct Ordered { bool on = false; int n = 1, me = 0; std::vector<int64_t> nch, goff; int64_t NG = 0, mxc = 0; };
#ifndef MC_HOSTEMU
static bool ordered_mode(const mc_ctx* c) { return c->comm && c->comm_size > 1 && !c->prm.reserved[2]; }
static int ordered_layout(mc_ctx* c, int64_t n_chunks, Ordered& o)
{
	o.on = true; o.n = c->comm_size; o.me = c->comm_rank;
	cudaStream_t s = c->stream;
	if (c->d_comm_small.reserve(8 * (o.n + 8))) return -1;
	long long* d_sz = c->d_comm_small.as<long long>();
	long long my = (long long)n_chunks; std::vector<long long> sz(o.n);
	if (dev_h2d(d_sz + o.n, &my, 8, s)) return -1;
	if (nccl_fail(ncclAllGather(d_sz + o.n, d_sz, 1, ncclInt64, c->comm, s), "ncclAllGather(chunk counts)")) return -1;
	if (dev_d2h(sz.data(), d_sz, 8 * o.n, s) || dev_sync(s)) return -1;
	o.nch.assign(o.n, 0); o.goff.assign(o.n + 1, 0); o.mxc = 1;
	for (int r = 0; r < o.n; r++) { o.nch[r] = sz[r]; o.goff[r + 1] = o.goff[r] + sz[r]; o.mxc = std::max(o.mxc, (int64_t)sz[r]); }
	o.NG = o.goff[o.n];
	return 0;
}
// all ranks' chunk sums, validity intervals and overflow flags after an attempt
static int ordered_chunks(mc_ctx* c, const Ordered& o, std::vector<mc_chunk_out>& hc, std::vector<int32_t>& lo, std::vector<int32_t>& hi, mc_u64* any_overflow)
{
	cudaStream_t s = c->stream;
	const bool dbg = getenv("MC_DEBUG") != nullptr;
	auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t0 = 0, t1 = 0, t2 = 0;
	if (dbg) { t0 = now(); t1 = t0; }
	const size_t cb = (size_t)o.mxc * sizeof(mc_chunk_out), ib = (size_t)o.mxc * 4, per = cb + 2 * ib + 8;
	if (c->d_comm_buf.reserve(per * o.n + 64) || c->h_comm_buf.reserve(per * o.n + 64)) return -1;
	uint8_t* d = c->d_comm_buf.as<uint8_t>();
	int bad = 0;
	nccl_api()->GroupStart();
	bad |= nccl_fail(ncclAllGather(c->d_chunk_out.p, d, cb, ncclUint8, c->comm, s), "ncclAllGather(chunk sums)");
	bad |= nccl_fail(ncclAllGather(c->d_chunk_lo.p, d + cb * o.n, ib, ncclUint8, c->comm, s), "ncclAllGather(chunk lo)");
	bad |= nccl_fail(ncclAllGather(c->d_chunk_hi.p, d + (cb + ib) * o.n, ib, ncclUint8, c->comm, s), "ncclAllGather(chunk hi)");
	bad |= nccl_fail(ncclAllGather(&c->d_stats.as<DevStats>()->overflow, d + (cb + 2 * ib) * o.n, 8, ncclUint8, c->comm, s), "ncclAllGather(overflow)");
	bad |= nccl_fail(nccl_api()->GroupEnd(), "ncclGroupEnd");
	if (bad) return -1;
	if (dbg) t2 = now();
	const double t3 = t2;
	if (dev_d2h(c->h_comm_buf.p, d, per * o.n, s) || dev_sync(s)) return -1;
	if (dbg) fprintf(stderr, "[mc] rank %d exchange: nccl enqueue %.3f ms, wait + d2h %.3f ms\n", o.me, t2 - t1, now() - t3);
	const uint8_t* h = c->h_comm_buf.as<uint8_t>();
	*any_overflow = 0;
	for (int r = 0; r < o.n; r++)
	{
		memcpy(hc.data() + o.goff[r], h + cb * r, (size_t)o.nch[r] * sizeof(mc_chunk_out));
		memcpy(lo.data() + o.goff[r], h + cb * o.n + ib * r, (size_t)o.nch[r] * 4);
		memcpy(hi.data() + o.goff[r], h + (cb + ib) * o.n + ib * r, (size_t)o.nch[r] * 4);
		mc_u64 f; memcpy(&f, h + (cb + 2 * ib) * o.n + 8 * r, 8); *any_overflow |= f;
	}
	return 0;
}
// gathers variable-length byte records of every rank into every rank: sizes first, then ONE padded all-gather
static int allgather_bytes(mc_ctx* c, ncclComm_t comm, const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t> >& all)
{
	const int n = c->comm_size; cudaStream_t s = c->stream;
	if (c->d_comm_small.reserve(8 * (n + 8))) return -1;
	long long* d_sz = c->d_comm_small.as<long long>();
	long long my = (long long)mine.size();
	std::vector<long long> sz(n);
	if (dev_h2d(d_sz + n, &my, 8, s)) return -1;
	if (nccl_fail(ncclAllGather(d_sz + n, d_sz, 1, ncclInt64, comm, s), "ncclAllGather(sizes)")) return -1;
	if (dev_d2h(sz.data(), d_sz, 8 * n, s) || dev_sync(s)) return -1;
	long long mx = 16; for (int r = 0; r < n; r++) mx = std::max(mx, sz[r]);
	mx = (mx + 15) & ~15ll;
	if (c->d_comm_buf.reserve((size_t)mx * (n + 1)) || c->h_comm_buf.reserve((size_t)mx * n)) return -1;
	uint8_t* d_all = c->d_comm_buf.as<uint8_t>(); uint8_t* d_mine = d_all + (size_t)mx * n;
	if (dev_h2d(d_mine, mine.data(), mine.size(), s)) return -1;
	if (nccl_fail(ncclAllGather(d_mine, d_all, (size_t)mx, ncclUint8, comm, s), "ncclAllGather(records)")) return -1;
	if (dev_d2h(c->h_comm_buf.p, d_all, (size_t)mx * n, s) || dev_sync(s)) return -1;
	all.assign(n, std::vector<uint8_t>());
	for (int r = 0; r < n; r++) all[r].assign(c->h_comm_buf.as<uint8_t>() + (size_t)mx * r, c->h_comm_buf.as<uint8_t>() + (size_t)mx * r + sz[r]);
	return 0;
}

// n 64-bit words of every rank, rank after rank, on the host of every rank (scalars of the distributed read-out)
static int gather_words(mc_ctx* c, const mc_u64* mine, int n, std::vector<mc_u64>& all)
{
	const int N = c->comm_size; cudaStream_t s = c->stream;
	if (c->d_comm_small.reserve(8 * (size_t)(n * (N + 1) + 8))) return -1;
	mc_u64* d = c->d_comm_small.as<mc_u64>();
	if (dev_h2d(d + (size_t)n * N, mine, 8 * (size_t)n, s)) return -1;
	if (nccl_fail(ncclAllGather(d + (size_t)n * N, d, (size_t)n, ncclUint64, c->comm, s), "ncclAllGather(scalars)")) return -1;
	all.assign((size_t)n * N, 0);
	if (dev_d2h(all.data(), d, 8 * (size_t)n * N, s) || dev_sync(s)) return -1;
	return 0;
}

#else
static bool ordered_mode(const mc_ctx*) { return false; }
#endif

// ---- the profile stage of a batch: UpdateProfile / UpdateMultiHitCount (reference src/AlignmentProfile.cpp:41-271) for the reads
// whose candidates, fragments and alignment strings lie in the batch's device arenas (`a`).  Runs at the end of run_batch, or
// later through mc_update_profile_last when the context defers it (mc_defer_profile).
static int profile_stage(mc_ctx* c, PipeArgs& a, int64_t n, const Ordered& od, int64_t frag_cap, const Bumps& hb)
{
	const mc_stream_t s = c->stream;
	Bumps* db = c->d_bumps.as<Bumps>();
	int64_t* h_small = c->h_small.as<int64_t>();
	int bad = dev_zero(&db->key, 8, s);      // the gate-key cursor (a deferred call finds the discordant-pair list's count in it)
	{
		PersistBumps pb;
		bad |= dev_d2h(&pb, c->d_pbump.p, sizeof(pb), s) || dev_sync(s);
		const int64_t need_bp = (int64_t)pb.bp + 2 * n, need_ind = (int64_t)pb.ind + (int64_t)hb.frag + 16, need_seq = (int64_t)pb.ind_seq + (int64_t)hb.aln + 16;
		bad |= c->d_bp.grow_keep(need_bp * 8, pb.bp * 8, s) || c->d_ind.grow_keep(need_ind * sizeof(mc_indel_rec), pb.ind * sizeof(mc_indel_rec), s);
		bad |= c->d_ind_seq.grow_keep(need_seq, pb.ind_seq, s);
		bad |= c->d_keys.reserve((n + 1) * 8) || c->d_keys_tmp.reserve((n + 1) * 8) || c->d_sort.reserve(device_sort_scratch_bytes(n));
		if (bad) return MC_ERR_CUDA;
		c->bp_cap = c->d_bp.cap / 8; c->ind_cap = c->d_ind.cap / sizeof(mc_indel_rec); c->ind_seq_cap = c->d_ind_seq.cap;
		if (c->ind_seq_cap >= 0x7fffffffll) { mc_set_error("indel sequence arena exceeds 2 GiB; call mc_profile_indels() earlier"); return MC_ERR_OVERFLOW; }
		PersistBumps* dpb = c->d_pbump.as<PersistBumps>();
		ProfArgs q; memset(&q, 0, sizeof(q));
		q.keys = c->d_keys.as<uint64_t>(); q.key_bump = &db->key; q.accept = c->d_accept.as<uint8_t>(); q.rnp = c->d_rnp.as<int32_t>();
		q.bp_pos = c->d_bp.as<int64_t>(); q.bp_bump = &dpb->bp; q.bp_cap = c->bp_cap;
		q.ind = c->d_ind.as<mc_indel_rec>(); q.ind_bump = &dpb->ind; q.ind_cap = c->ind_cap;
		q.ind_seq = c->d_ind_seq.as<uint8_t>(); q.ind_seq_bump = &dpb->ind_seq; q.ind_seq_cap = c->ind_seq_cap;
		launch_profkey(a, q, n, s);
		bad |= dev_d2h(h_small, &db->key, 8, s) || dev_sync(s);
		if (bad) return MC_ERR_CUDA;
		q.n_keys = h_small[0];
		device_sort_u64(q.keys, c->d_keys_tmp.as<uint64_t>(), q.n_keys, c->d_sort.p, c->d_sort.cap, s);
#ifndef MC_HOSTEMU
		std::vector<long long> gl_n; long long gl_mx = 0; const uint64_t* gl_all = nullptr;
		const uint8_t* gd_all = nullptr; size_t gd_pitch = 0;
		if (od.on)
		{
			// per-start candidate counts of every rank (heads of the sorted key runs, capped at 15).  Small genomes: one byte
			// per column, all-gathered and applied by streaming kernels; large ones: (start, count) lists, all-gathered and
			// applied list by list.  The choice depends only on sizes every rank knows (G and the key counts).
			long long* d_sz = c->d_comm_small.as<long long>();
			long long my_keys = q.n_keys;
			if (dev_h2d(d_sz + od.n, &my_keys, 8, s)) return MC_ERR_CUDA;
			if (nccl_fail(ncclAllGather(d_sz + od.n, d_sz, 1, ncclInt64, c->comm, s), "ncclAllGather(key counts)")) return MC_ERR_NCCL;
			gl_n.resize(od.n);
			if (dev_d2h(gl_n.data(), d_sz, 8 * od.n, s) || dev_sync(s)) return MC_ERR_CUDA;
			long long keys_mx = 0; for (int r = 0; r < od.n; r++) keys_mx = std::max(keys_mx, gl_n[r]);
			if (keys_mx == 0) {}
			else if ((long long)c->G <= 8 * keys_mx)
			{
				gd_pitch = ((size_t)c->G + 255) & ~(size_t)255;
				if (c->d_glist.reserve(gd_pitch) || c->d_glist_all.reserve(gd_pitch * od.n) || dev_zero(c->d_glist.p, gd_pitch, s)) return MC_ERR_CUDA;
				launch_gatedense_fill(a, q, q.n_keys, c->d_glist.as<uint8_t>(), s);
				if (nccl_fail(ncclAllGather(c->d_glist.p, c->d_glist_all.p, gd_pitch, ncclUint8, c->comm, s), "ncclAllGather(gate counts)")) return MC_ERR_NCCL;
				gd_all = c->d_glist_all.as<uint8_t>();
				launch_gatedense_apply(a, c->G, gd_all, gd_pitch, 0, od.me, s);
			}
			else
			{
				if (c->d_glist.reserve((size_t)(q.n_keys + 2) * 8) || dev_zero(&db->rwin, 8, s)) return MC_ERR_CUDA;   // the window cursor is free: reuse it
				launch_gatecnt(a, q, q.n_keys, c->d_glist.as<uint64_t>(), &db->rwin, s);
				if (nccl_fail(ncclAllGather(&db->rwin, d_sz, 1, ncclInt64, c->comm, s), "ncclAllGather(gate list sizes)")) return MC_ERR_NCCL;
				if (dev_d2h(gl_n.data(), d_sz, 8 * od.n, s) || dev_sync(s)) return MC_ERR_CUDA;
				for (int r = 0; r < od.n; r++) gl_mx = std::max(gl_mx, gl_n[r]);
				gl_mx = (gl_mx + 1) & ~1ll;
				if (gl_mx)
				{
					if (c->d_glist.grow_keep((size_t)gl_mx * 8, (size_t)gl_n[od.me] * 8, s) || c->d_glist_all.reserve((size_t)gl_mx * 8 * od.n)) return MC_ERR_CUDA;
					if (nccl_fail(ncclAllGather(c->d_glist.p, c->d_glist_all.p, (size_t)gl_mx * 8, ncclUint8, c->comm, s), "ncclAllGather(gate lists)")) return MC_ERR_NCCL;
					gl_all = c->d_glist_all.as<uint64_t>();
					for (int r = 0; r < od.me; r++) launch_gateadd(a, gl_n[r], gl_all + (size_t)gl_mx * r, s);
				}
			}
		}
#endif
		launch_gate(a, q, q.n_keys, s);
		launch_gateupd(a, q, q.n_keys, s);
#ifndef MC_HOSTEMU
		if (od.on && gl_all) for (int r = od.me + 1; r < od.n; r++) launch_gateadd(a, gl_n[r], gl_all + (size_t)gl_mx * r, s);
		if (od.on && gd_all) launch_gatedense_apply(a, c->G, gd_all, gd_pitch, od.me + 1, od.n, s);
#endif
		bad |= dev_zero(&db->ptask, 8, s);            // the piece list of the alignment stage is free again: reuse it
		launch_scatter(a, q, n, s);
		launch_profpiece(a, q, frag_cap, s);
		}
	return bad ? MC_ERR_CUDA : MC_OK;
}

// ---- the batch controller -----------------------------------------------------------------------------
static int run_batch(mc_ctx* c, Staged& st, mc_batch_out* out, bool prep_needed)
{
	const mc_stream_t s = c->stream;
	const int64_t n = st.n_reads;
	const bool paired = c->prm.paired != 0;
	const int64_t n_pairs = paired ? n / 2 : 0;
	const int64_t n_chunks = (n + MC_CHUNK_READS - 1) / MC_CHUNK_READS;
	memset(out, 0, sizeof(*out));
	c->last_n = 0;
	// the reference cuts the library into 200-read chunks from its start; a batch that ends inside a chunk can only be the last
	if (n > 0 && c->library_closed && !ordered_mode(c)) { mc_set_error("mc_map_batch: the previous batch was not a multiple of %d reads, which ends the library (mc_reset starts a new one)", MC_CHUNK_READS); return MC_ERR_ARG; }
	if (n % MC_CHUNK_READS) c->library_closed = true;
	if (n == 0 && ordered_mode(c)) { mc_set_error("mc_map_batch: with the ordered multi-GPU exchange every rank has to pass reads in every call (mc_comm_init)"); return MC_ERR_ARG; }
	if (n == 0) return MC_OK;

	PipeArgs a; memset(&a, 0, sizeof(a));
	a.ix = c->ix;
	a.pr.paired = c->prm.paired; a.pr.alg_ksw2 = c->prm.alg_ksw2; a.pr.max_pos_diff = c->prm.max_pos_diff; a.pr.max_clip = c->prm.max_clip;
	a.pr.max_dup = c->prm.max_dup; a.pr.update_profile = c->prm.update_profile; a.pr.max_mismatch_rate = c->prm.max_mismatch_rate;
	a.st = c->d_stats.as<DevStats>(); a.n_reads = n; a.seq = st.seq.as<uint8_t>() - st.base; a.roff = st.roff.as<int64_t>(); a.seed_off = st.seed_off.as<int64_t>();
	a.first_read = c->tot.total_reads;
	a.n_slots = st.n_slots;

	int bad = 0;
	bad |= c->d_slot_freq.reserve(st.n_slots * 4) || c->d_seeds.reserve(st.n_slots * sizeof(Seed)) || c->d_slot_loc.reserve((st.n_slots + 1) * 8);
	bad |= c->d_scan.reserve(device_scan_scratch_bytes(st.n_slots));
	bad |= c->d_cand_off.reserve((n + 1) * 4) || c->d_rflag.reserve(n + 1) || c->d_npair.reserve(n * 4) || c->d_ncand0.reserve(n * 4) || c->d_ncand.reserve(n * 4) || c->d_rsum.reserve(n * sizeof(ReadSum));
	// several GPUs on one library: the chunk grid of the controller below is the global one, this rank owns [g0, g0 + n_chunks)
	Ordered od;
#ifndef MC_HOSTEMU
	if (ordered_mode(c) && ordered_layout(c, n_chunks, od)) return MC_ERR_NCCL;
#endif
	const int64_t NG = od.on ? od.NG : n_chunks, g0 = od.on ? od.goff[od.me] : 0, mxc = od.on ? od.mxc : n_chunks;
	bad |= c->d_est.reserve(n_chunks * 4) || c->d_active.reserve(n_chunks) || c->d_chunk_out.reserve(mxc * sizeof(mc_chunk_out));
	bad |= c->d_chunk_lo.reserve(mxc * 4) || c->d_chunk_hi.reserve(mxc * 4);
	bad |= c->d_pair_flag.reserve((n_pairs + 1) * 4) || c->d_est_lo.reserve((n_pairs + 1) * 4) || c->d_est_hi.reserve((n_pairs + 1) * 4);
	bad |= c->d_pair_out.reserve((n_pairs + 1) * sizeof(mc_pair_out)) || c->d_rtask.reserve((n_pairs + 1) * 4 * 4) || c->d_rw_beg.reserve((n_pairs + 1) * 4) || c->d_accept.reserve(n + 1) || c->d_rnp.reserve((n + 1) * 4) || c->d_read_redo.reserve(n + 1);
	bad |= c->h_chunk.reserve(n_chunks * sizeof(mc_chunk_out)) || c->h_chunk_lo.reserve(n_chunks * 4) || c->h_chunk_hi.reserve(n_chunks * 4);
	if (bad) return MC_ERR_CUDA;
	a.slot_freq = c->d_slot_freq.as<uint32_t>(); a.seeds = c->d_seeds.as<Seed>(); a.slot_loc = c->d_slot_loc.as<int64_t>();
	a.cand_off = c->d_cand_off.as<int32_t>(); a.rflag = c->d_rflag.as<uint8_t>(); a.npair = c->d_npair.as<int32_t>(); a.ncand0 = c->d_ncand0.as<int32_t>(); a.ncand = c->d_ncand.as<int32_t>(); a.rsum = c->d_rsum.as<ReadSum>();
	a.est = c->d_est.as<int32_t>(); a.active = c->d_active.as<uint8_t>(); a.chunk_out = c->d_chunk_out.as<mc_chunk_out>();
	a.chunk_lo = c->d_chunk_lo.as<int32_t>(); a.chunk_hi = c->d_chunk_hi.as<int32_t>();
	a.pair_flag = c->d_pair_flag.as<int32_t>(); a.est_lo = c->d_est_lo.as<int32_t>(); a.est_hi = c->d_est_hi.as<int32_t>();
	a.pair_out = c->d_pair_out.as<mc_pair_out>(); a.rtask = c->d_rtask.as<int32_t>(); a.rw_beg = c->d_rw_beg.as<int32_t>(); a.read_redo = c->d_read_redo.as<uint8_t>();
	const int64_t rtask_cap = (n_pairs + 1) * 4;
	Bumps* db = c->d_bumps.as<Bumps>();
	a.pair_bump = &db->pair; a.frag_bump = &db->frag; a.aln_bump = &db->aln; a.task_bump = &db->task; a.dpws_bump = &db->dpws; a.rtask_bump = &db->rtask; a.ptask_bump = &db->ptask; a.rwin_bump = &db->rwin; a.rwin_begin = 0;
	a.prof.base16 = c->d_base16.as<uint32_t>(); a.prof.sdiff = c->d_sdiff.as<int32_t>(); a.prof.cdiff = c->d_cdiff.as<int32_t>(); a.prof.mdiff = c->d_mdiff.as<int32_t>();
	a.prof.rcount = c->d_rcount.as<uint8_t>();

	// ---- seeding ----
	ev_record(&c->ev[EV_H2D], s);
	bad |= dev_zero(c->d_slot_freq.p, st.n_slots * 4, s) || dev_zero(c->d_stats.p, sizeof(DevStats) - 2 * sizeof(mc_u64), s);
	bad |= dev_d2d(&c->d_stats.as<DevStats>()->overflow, st.flag.p, 8, s);   // what staging found (a read longer than MC_MAX_RLEN)
	ev_record(&c->ev[EV_SEED0], s);
	if (st.n_pieces > 1)
	{
#ifndef MC_HOSTEMU
		for (int p = 0; p < st.n_pieces; p++)
		{
			const int64_t r0 = p ? st.piece_end[p - 1] : 0, r1 = st.piece_end[p];
			if (cuda_fail(cudaStreamWaitEvent(s, c->ev_piece[p], 0), "cudaStreamWaitEvent")) return MC_ERR_CUDA;
			if (prep_needed) launch_prep(a, r0, r1, s);
			launch_seed(a, r0, r1, s);
		}
#endif
	}
	else
	{
		if (prep_needed) launch_prep(a, 0, n, s);
		launch_seed(a, 0, n, s);
	}
	ev_record(&c->ev[EV_SEED], s);
	device_scan_u32(a.slot_freq, c->d_slot_loc.as<int64_t>(), st.n_slots, c->d_scan.as<int64_t>(), s);
	int64_t* h_small = c->h_small.as<int64_t>();
	bad |= dev_d2h(h_small, c->d_slot_loc.as<int64_t>() + st.n_slots, 8, s) || dev_sync(s);
	if (bad) return MC_ERR_CUDA;
	const int64_t n_locs = h_small[0];
	a.n_locs = n_locs;
	if (getenv("MC_DEBUG")) fprintf(stderr, "[mc] batch: %lld reads, %lld bytes, %lld seed slots, %lld seed locations\n", (long long)n, (long long)st.n_bytes, (long long)st.n_slots, (long long)n_locs);

	// ---- locate + cluster ----
	const int64_t cand_total = (paired ? 2 : 1) * n_locs + 1;
	for (;;) // arenas may have to grow (overflow flag) -> retry from here
	{
		const int64_t pair_cap = n_locs + c->rescue_cap;
		const int64_t frag_cap = (int64_t)(c->frag_factor * (double)(n_locs + 1024)) + n;
		const int64_t aln_cap = (int64_t)(c->aln_factor * (double)st.n_bytes) + (1 << 20);
		const int64_t dpws_cap = (int64_t)(c->dpws_factor * (double)st.n_bytes) + (8 << 20);
		const int64_t task_cap = (int64_t)(c->task_factor * (double)n) + 1024;
		if (frag_cap >= 0x7fffffffll || aln_cap >= 0x7fffffffll || cand_total >= 0x7fffffffll) { mc_set_error("mc_map_batch: batch too large for 32-bit arena offsets; split it"); return MC_ERR_ARG; }
		bad |= c->d_pairs.reserve(pair_cap * sizeof(SPair)) || c->d_rwin.reserve(c->rescue_cap * sizeof(RWin));
		bad |= c->d_cands.reserve(cand_total * sizeof(Cand)) || c->d_cscore.reserve(cand_total * 4) || c->d_cpaired.reserve(cand_total * 4);
		bad |= c->d_corient.reserve(cand_total * 4) || c->d_cfrag.reserve(cand_total * 4) || c->d_cnfrag.reserve(cand_total * 4) || c->d_ctmp.reserve(cand_total * 4);
		bad |= c->d_ptask.reserve(frag_cap * 4) || c->d_frags.reserve(frag_cap * sizeof(mc_frag_out)) || c->d_aln.reserve(aln_cap) || c->d_tasks.reserve(task_cap * sizeof(DpTask)) || c->d_dpws.reserve(dpws_cap);
		if (bad) return MC_ERR_CUDA;
		a.pairs = c->d_pairs.as<SPair>(); a.pair_cap = pair_cap; a.rwin = c->d_rwin.as<RWin>(); a.rwin_cap = c->rescue_cap;
		a.cands = c->d_cands.as<Cand>(); a.cscore = c->d_cscore.as<int32_t>(); a.cpaired = c->d_cpaired.as<int32_t>(); a.corient = c->d_corient.as<int32_t>();
		a.cfrag = c->d_cfrag.as<int32_t>(); a.cnfrag = c->d_cnfrag.as<int32_t>(); a.ctmp = c->d_ctmp.as<int32_t>();
		a.ptask = c->d_ptask.as<int32_t>(); a.frags = c->d_frags.as<mc_frag_out>(); a.frag_cap = frag_cap; a.aln = c->d_aln.as<uint8_t>(); a.aln_cap = aln_cap;
		a.tasks = c->d_tasks.as<DpTask>(); a.task_cap = task_cap; a.dpws = c->d_dpws.as<uint8_t>(); a.dpws_cap = dpws_cap;

		launch_expand(a, st.n_slots, s);
		ev_record(&c->ev[EV_LOC0], s);
		launch_locate(a, n_locs, s);
		ev_record(&c->ev[EV_LOCATE], s);
		launch_cluster(a, n, s);
		ev_record(&c->ev[EV_CLUSTER], s);

		// ---- speculative pairing / alignment with avgDist verification ----
		std::vector<int32_t> est(NG, (int32_t)(c->tot.avg_dist * 1.5));
		std::vector<uint8_t> active(NG, 1), computed(NG, 0), ever(NG, 0);
		std::vector<mc_chunk_out> g_hc; std::vector<int32_t> g_lo, g_hi;
		if (od.on) { g_hc.resize(NG); g_lo.resize(NG); g_hi.resize(NG); }
		// avgDist stays at its initial value until more than 1000 pairs have been seen (src/ReadMapping.cpp:539) and then
		// jumps: while warming up only the chunks that can still use the initial value are speculated on - plus a few hundred
		// more, whose sums (nearly independent of the value) let the walk predict where the trajectory settles
		if (paired && c->tot.total_paired <= 1000 && !c->freeze_avg_dist)
			for (int64_t k = (1000 - c->tot.total_paired) / (MC_CHUNK_READS / 2) + 2 + 256; k < NG; k++) active[k] = 0;
		Bumps hb; memset(&hb, 0, sizeof(hb)); hb.pair = (mc_u64)n_locs;
		bad |= dev_push(c, db, &hb, sizeof(hb), s) || dev_zero(c->d_pair_flag.p, (n_pairs + 1) * 4, s);
		mc_totals run = c->tot;
		c->chunks_final.assign(NG, mc_chunk_out());
		int64_t first_open = 0;
		int replays = 0; bool overflow = false, first_attempt = true;
		mc_u64 ovf_all = 0;                    // ordered exchange: OR of every rank's overflow word, so that all ranks leave together
		Bumps* hbp = (Bumps*)(h_small + 64);   // pinned copy of the arena cursors, refreshed at the end of every attempt
		memset(hbp, 0, sizeof(Bumps));
		ev_record(&c->ev[EV_PAIR0], s);
		while (first_open < NG)
		{
			const bool dbg_att = getenv("MC_DEBUG") != nullptr;
			const auto att_t0 = std::chrono::steady_clock::now();
			// one attempt: no host round trip inside it, the task lists are consumed from their device-side cursors
			a.rtask_begin = (int64_t)hbp->rtask; a.task_begin = (int64_t)hbp->task; a.ptask_begin = (int64_t)hbp->ptask;
			if (a.rtask_begin + n_pairs > rtask_cap) { mc_set_error("mc_map_batch: too many speculation replays in one batch"); return MC_ERR_OVERFLOW; }
			bad |= dev_push(c, c->d_est.p, est.data() + g0, n_chunks * 4, s) || dev_push(c, c->d_active.p, active.data() + g0, n_chunks, s);
			if (paired) { bad |= dev_zero(&db->rwin, 8, s); launch_pair(a, n_pairs, s); launch_rescue(a, n_pairs, s); } else launch_single(a, n, s);
			if (first_attempt) ev_record(&c->ev[EV_PAIR1], s);
			first_attempt = false;
			launch_alnprep(a, n, s);
			launch_piece(a, frag_cap, s);
			launch_dp(a, task_cap - a.task_begin, s);
			launch_alnfin(a, n, s);
			if (paired) launch_pairstat(a, n_pairs, s);
			launch_chunkstat(a, n_chunks, s);
			bad |= dev_d2h(hbp, db, sizeof(Bumps), s);
			DevStats* hst = (DevStats*)(h_small + 8);
			bad |= dev_d2h(c->h_chunk.p, c->d_chunk_out.p, n_chunks * sizeof(mc_chunk_out), s) || dev_d2h(c->h_chunk_lo.p, c->d_chunk_lo.p, n_chunks * 4, s);
			bad |= dev_d2h(c->h_chunk_hi.p, c->d_chunk_hi.p, n_chunks * 4, s) || dev_d2h(hst, c->d_stats.p, sizeof(DevStats), s) || dev_sync(s);
			if (bad) return MC_ERR_CUDA;
			const mc_chunk_out* hc = c->h_chunk.as<mc_chunk_out>(); const int32_t* lo = c->h_chunk_lo.as<int32_t>(); const int32_t* hi = c->h_chunk_hi.as<int32_t>();
#ifndef MC_HOSTEMU
			if (od.on)
			{
				mc_u64 any = 0;
				if (ordered_chunks(c, od, g_hc, g_lo, g_hi, &any)) return MC_ERR_NCCL;
				hc = g_hc.data(); lo = g_lo.data(); hi = g_hi.data();
				if (any) { overflow = true; ovf_all = any; break; }      // every rank repeats the batch (each grows only what it ran out of)
			}
#endif
			if (hst->overflow) { overflow = true; break; }
			if (dbg_att)
			{
				int64_t na = 0, nl = 0; for (int64_t k = 0; k < NG; k++) { na += active[k]; if (k >= g0 && k < g0 + n_chunks) nl += active[k]; }
				fprintf(stderr, "[mc] rank %d attempt: %lld active chunks (%lld mine) of %lld, first open %lld, %.3f ms; rescue pairs %lld windows %lld, pieces %lld, fills %lld\n", od.on ? od.me : 0, (long long)na, (long long)nl, (long long)NG, (long long)first_open,
				        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - att_t0).count(), (long long)hbp->rtask - (long long)a.rtask_begin, (long long)hbp->rwin, (long long)hbp->ptask - (long long)a.ptask_begin, (long long)hbp->task - (long long)a.task_begin);
			}
			for (int64_t k = 0; k < NG; k++) if (active[k]) computed[k] = ever[k] = 1;
			// walk the chunks in file order (reference src/ReadMapping.cpp:537-539)
			std::fill(active.begin(), active.end(), 0);
			while (first_open < NG)
			{
				const int64_t k = first_open;
				const int32_t true_est = (int32_t)(run.avg_dist * 1.5);
				if (!computed[k] || (paired && !(lo[k] <= true_est && true_est <= hi[k])))
				{
					// (re)run everything from here on: this chunk with the value it really sees, the later ones with the values
					// the trajectory is going to take if their sums stay what the previous attempt found (they nearly always do:
					// only pairs at the edge of the distance window move) - the walk verifies every one of them again
					if (computed[k]) replays++;
					mc_totals pred = run;
					for (int64_t j = k; j < NG; j++)
					{
						est[j] = (int32_t)(pred.avg_dist * 1.5); active[j] = 1; computed[j] = 0;
						if (!ever[j]) continue;
						pred.total_paired += hc[j].paired; pred.total_distance += hc[j].dist_sum;
						if (paired && pred.total_paired > 1000 && !c->freeze_avg_dist) pred.avg_dist = (uint32_t)(int)(1. * pred.total_distance / pred.total_paired + .5);
					}
					break;
				}
				mc_chunk_out ck = hc[k]; ck.est_distance = paired ? true_est : 0;
				c->chunks_final[k] = ck;
				run.total_reads += ck.n_reads; run.total_mapped += ck.mapped; run.total_paired += ck.paired;
				run.total_distance += ck.dist_sum; run.read_length_sum += ck.len_sum;
				if (paired && run.total_paired > 1000 && !c->freeze_avg_dist) run.avg_dist = (uint32_t)(int)(1. * run.total_distance / run.total_paired + .5);
				first_open++;
			}
		}
		if (overflow)
		{
			const mc_u64 ovf = ((DevStats*)(h_small + 8))->overflow;
			const mc_u64 fatal = ovf | ovf_all;       // a fatal bit on ANY rank ends the collective call on every rank with the same code
			if ((ovf >> 0) & 0xFF) c->rescue_cap *= 4;
			if ((ovf >> 8) & 0xFF) c->frag_factor *= 2;
			if ((ovf >> 16) & 0xFF) c->aln_factor *= 2;
			if ((ovf >> 24) & 0xFF) c->task_factor *= 2;
			if ((ovf >> 32) & 0xFF) c->dpws_factor *= 4;
			if ((fatal >> 56) & 0xFF) { mc_set_error("mc_map_batch: a read is longer than %d bases (or the offsets decrease)%s", MC_MAX_RLEN, ((ovf >> 56) & 0xFF) ? "" : " on another rank"); dev_zero(c->d_stats.p, sizeof(DevStats), s); return MC_ERR_ARG; }
			if ((fatal >> 40) & 0xFF) { mc_set_error("mc_map_batch: internal error: candidate table overflow%s", ((ovf >> 40) & 0xFF) ? "" : " on another rank"); return MC_ERR_OVERFLOW; }
			if (c->frag_factor > 4096 || c->aln_factor > 4096 || c->task_factor > 4096 || c->dpws_factor > 65536) { mc_set_error("mc_map_batch: arena overflow persists"); return MC_ERR_OVERFLOW; }
			if (getenv("MC_DEBUG")) fprintf(stderr, "[mc] arena overflow %llx: now frag x%.0f aln x%.0f task x%.0f dpws x%.0f rescue %lld\n", (unsigned long long)ovf, c->frag_factor, c->aln_factor, c->task_factor, c->dpws_factor, (long long)c->rescue_cap);
			bad |= dev_zero(&c->d_stats.as<DevStats>()->locate_blocks, sizeof(DevStats) - 3 * sizeof(mc_u64), s); // keep only the seed kernel's counters
			continue;
		}
		ev_record(&c->ev[EV_ALN1], s);
		out->replays = replays;

		// ---- profile ----
		bad |= dev_d2h(&hb, db, sizeof(hb), s) || dev_sync(s);
		if (bad) return MC_ERR_CUDA;
		ev_record(&c->ev[EV_PROF0], s);
		c->last_a = a; c->last_frag_cap = frag_cap; c->last_hb = hb; c->last_profiled = false;
		if (c->prm.update_profile && !c->defer_profile)
		{
			if (int prc = profile_stage(c, a, n, od, frag_cap, hb)) return prc;
			c->last_profiled = true;
		}
		ev_record(&c->ev[EV_PROF1], s);

		// ---- results back to the host ----
		const bool want_pairs = paired && (c->prm.want_alignments || c->prm.reserved[0]);
		if (want_pairs) { bad |= c->h_pairs.reserve((n_pairs + 1) * sizeof(mc_pair_out)); bad |= dev_d2h(c->h_pairs.p, c->d_pair_out.p, n_pairs * sizeof(mc_pair_out), s); }
		// the few pairs that are neither proper nor unanchored go to the host for the in-order SV-site bookkeeping
		int64_t disc_cap = 0;
		if (paired && c->prm.update_profile)
		{
			disc_cap = n_pairs;
			bad |= c->d_disc.reserve((size_t)disc_cap * sizeof(DiscRec) + 64) || dev_zero(&db->key, 8, s);   // the gate-key cursor is free again
			launch_disclist(a, n_pairs, c->d_disc.as<DiscRec>(), &db->key, disc_cap, s);
			bad |= dev_d2h(h_small + 32, &db->key, 8, s) || dev_sync(s);
			if (bad) return MC_ERR_CUDA;
			const int64_t nd = std::min(h_small[32], disc_cap);
			bad |= c->h_disc.reserve((size_t)nd * sizeof(DiscRec) + 64) || dev_d2h(c->h_disc.p, c->d_disc.p, (size_t)nd * sizeof(DiscRec), s);
			h_small[33] = nd;
		}
		if (c->prm.want_alignments)
		{
			const size_t cb = (size_t)cand_total * 4;
			bad |= c->h_reads.reserve(n * (sizeof(ReadSum) + 4)) || c->h_cands.reserve(cb * 5) || c->h_frags.reserve((size_t)hb.frag * sizeof(mc_frag_out)) || c->h_aln.reserve((size_t)hb.aln + 16);
			if (bad) return MC_ERR_CUDA;
			bad |= dev_d2h(c->h_reads.p, c->d_rsum.p, n * sizeof(ReadSum), s) || dev_d2h(c->h_reads.as<uint8_t>() + n * sizeof(ReadSum), c->d_ncand.p, n * 4, s);
			uint8_t* hcb = c->h_cands.as<uint8_t>();
			bad |= dev_d2h(hcb, c->d_cscore.p, cb, s) || dev_d2h(hcb + cb, c->d_cpaired.p, cb, s) || dev_d2h(hcb + 2 * cb, c->d_corient.p, cb, s);
			bad |= dev_d2h(hcb + 3 * cb, c->d_cfrag.p, cb, s) || dev_d2h(hcb + 4 * cb, c->d_cnfrag.p, cb, s);
			bad |= dev_d2h(c->h_frags.p, c->d_frags.p, (size_t)hb.frag * sizeof(mc_frag_out), s) || dev_d2h(c->h_aln.p, c->d_aln.p, (size_t)hb.aln, s);
		}
		DevStats* hst = (DevStats*)(h_small + 8);
		bad |= dev_d2h(hst, c->d_stats.p, sizeof(DevStats), s);
		ev_record(&c->ev[EV_D2H], s);
		bad |= dev_sync(s);
		if (bad) return MC_ERR_CUDA;
		if (hst->overflow) { mc_set_error("mc_map_batch: persistent profile arena overflow"); return MC_ERR_OVERFLOW; }

		if (c->prm.want_alignments)
		{
			const ReadSum* rs = c->h_reads.as<ReadSum>(); const int32_t* nc = (const int32_t*)(c->h_reads.as<uint8_t>() + n * sizeof(ReadSum));
			c->reads_out.resize(n);
			for (int64_t r = 0; r < n; r++)
			{
				mc_read_out ro; ro.score = rs[r].score; ro.sub_score = rs[r].sub_score; ro.best_idx = rs[r].best_idx; ro.n_cand = nc[r];
				ro.rlen = (int32_t)(st.h_roff[r + 1] - st.h_roff[r]); ro.cand_begin = 0;
				c->reads_out[r] = ro;
			}
		}

		// pair classification that the reference does in file order with thread-local state (src/ReadMapping.cpp:486-522)
		if (paired && c->prm.update_profile)
		{
			DiscRec* d = c->h_disc.as<DiscRec>(); const int64_t nd = h_small[33];
			std::sort(d, d + nd, [](const DiscRec& x, const DiscRec& y) { return x.pair < y.pair; });
			const int64_t G = c->G, twoG = 2 * c->G;
			// `keep` = the records are this rank's own: their sites are stored; other ranks' records only move the state along
			auto replay = [&](const DiscRec* d, int64_t nd, bool keep) {
				for (int64_t k = 0; k < nd; k++)
				{
					const mc_pair_out& q = d[k].v;
					if (q.gPos1 < G && q.gPos2 >= G)
					{
						c->discord_dist = llabs(twoG - q.gPos1 - q.gPos2);
						if (c->discord_dist > 1000 && c->discord_dist < 10000000) { c->discord_gpos = q.gPos1; if (keep) c->inv_sites.push_back({c->discord_gpos, c->discord_dist}); }
					}
					else if (q.gPos1 >= G && q.gPos2 < G)
					{
						c->discord_dist = llabs(twoG - q.gPos1 - q.gPos2);
						if (c->discord_dist > 1000 && c->discord_dist < 10000000) c->discord_gpos = q.gPos2;
						if (keep) c->inv_sites.push_back({c->discord_gpos, c->discord_dist}); // the reference pushes unconditionally here (:502)
					}
					else if (q.dist > 1000)
					{
						c->discord_dist = q.dist;
						if (q.gPos1 < G && q.gPos2 < G) { if (keep) { c->tnl_sites.push_back({q.gPos1, q.dist}); c->tnl_sites.push_back({q.gPos2, q.dist}); } c->discord_gpos = q.gPos2; }
						else if (q.gPos1 >= G && q.gPos2 >= G) { if (keep) { c->tnl_sites.push_back({twoG - q.gPos1, q.dist}); c->tnl_sites.push_back({twoG - q.gPos2, q.dist}); } c->discord_gpos = twoG - q.gPos2; }
					}
				}
			};
#ifndef MC_HOSTEMU
			if (od.on)
			{
				std::vector<uint8_t> mine((const uint8_t*)d, (const uint8_t*)(d + nd)); std::vector<std::vector<uint8_t> > all;
				if (allgather_bytes(c, c->comm, mine, all)) return MC_ERR_NCCL;
				for (int r = 0; r < od.n; r++) replay((const DiscRec*)all[r].data(), (int64_t)(all[r].size() / sizeof(DiscRec)), r == od.me);
			}
			else
#endif
				replay(d, nd, true);
		}

		c->tot = run;
		// stats
		c->stats.ms_h2d += ev_ms(&c->ev[EV_START], &c->ev[EV_H2D]);
		c->stats.ms_seed += ev_ms(&c->ev[EV_SEED0], &c->ev[EV_SEED]);     // the seed kernel alone
		c->stats.ms_locate += ev_ms(&c->ev[EV_LOC0], &c->ev[EV_LOCATE]);   // the locate kernel alone
		c->stats.ms_cluster += ev_ms(&c->ev[EV_LOCATE], &c->ev[EV_CLUSTER]);
		c->stats.ms_pair += ev_ms(&c->ev[EV_PAIR0], &c->ev[EV_PAIR1]);
		c->stats.ms_align += ev_ms(&c->ev[EV_PAIR1], &c->ev[EV_ALN1]);
		c->stats.ms_profile += ev_ms(&c->ev[EV_PROF0], &c->ev[EV_PROF1]);
		c->stats.ms_d2h += ev_ms(&c->ev[EV_PROF1], &c->ev[EV_D2H]);
		c->stats.ms_total += ev_ms(&c->ev[EV_START], &c->ev[EV_D2H]);
		c->stats.seed_blocks += hst->seed_blocks; c->stats.locate_blocks += hst->locate_blocks + hst->seed_locate_blocks; c->stats.sa_reads += hst->sa_reads + hst->seed_sa_reads;
		c->stats.dp_cells += hst->dp_cells; c->stats.dp_tasks += hst->dp_tasks; c->stats.profile_columns += hst->profile_columns; c->stats.profile_atomics += hst->profile_atomics;
		c->dstats_last = *hst;

		out->n_reads = n; out->n_pairs = n_pairs; out->n_chunks = n_chunks;
		out->pairs = want_pairs ? c->h_pairs.as<mc_pair_out>() : nullptr; if (!want_pairs) out->n_pairs = 0;
		out->chunks = c->chunks_final.data() + g0;
		if (c->prm.want_alignments)
		{
			// rebuild the per-read candidate table from the device's structure-of-arrays
			const size_t ct = (size_t)cand_total; const int32_t* hc = c->h_cands.as<int32_t>();
			const int32_t *sc = hc, *pi = hc + ct, *ori = hc + 2 * ct, *fb = hc + 3 * ct, *nf = hc + 4 * ct;
			// the candidate slice of every read inside the device arena
			std::vector<int32_t> cand_off(n + 1);
			bad |= dev_d2h(cand_off.data(), c->d_cand_off.p, n * 4, s) || dev_sync(s);
			if (bad) return MC_ERR_CUDA;
			c->cands_out.clear();
			for (int64_t r = 0; r < n; r++)
			{
				const int64_t co = cand_off[r];
				mc_read_out& ro = c->reads_out[r];
				ro.cand_begin = (int32_t)c->cands_out.size();
				for (int k = 0; k < ro.n_cand; k++)
				{
					mc_cand_out o; o.score = sc[co + k]; o.orientation = o.score > 0 ? ori[co + k] : -1; o.paired_idx = pi[co + k];
					o.frag_begin = o.score > 0 ? fb[co + k] : 0; o.n_frag = o.score > 0 ? nf[co + k] : 0; o.pad = 0;
					c->cands_out.push_back(o);
				}
			}
			out->reads = c->reads_out.data(); out->cands = c->cands_out.data(); out->n_cands = (int64_t)c->cands_out.size();
			out->frags = c->h_frags.as<mc_frag_out>(); out->n_frags = (int64_t)hb.frag; out->aln = c->h_aln.as<uint8_t>(); out->n_aln_bytes = (int64_t)hb.aln;
		}
		c->last_n = n; c->last_roff = st.roff.as<int64_t>();
		return MC_OK;
	}
}

// EvaluateMAPQ (src/SamReport.cpp:86-101) mixes float and double arithmetic with log(): tabulated with the host's libm for
// the only cases that reach the formula (score - sub_score = 1..5), so the device needs no transcendental
static int sam_mapq_table(mc_ctx* c, DBuf& d_tab, mc_stream_t s)
{
	if (d_tab.cap) return 0;
	std::vector<uint8_t> tab((size_t)5 * (MC_MAX_RLEN + 1), 0);
	for (int d = 1; d <= 5; d++)
		for (int score = d; score <= MC_MAX_RLEN; score++)
		{
			int mapq = (int)(30 * (1 - (float)d / score) * log(score) + 0.4999);
			if (mapq > 60) mapq = 60;
			tab[(size_t)(d - 1) * (MC_MAX_RLEN + 1) + score] = (uint8_t)mapq;
		}
	return d_tab.reserve(tab.size()) || dev_h2d(d_tab.p, tab.data(), tab.size(), s) || dev_sync(s);
}
static void sam_args_of(mc_ctx* c, SamArgs& a, int64_t n, DBuf& d_tab)
{
	memset(&a, 0, sizeof(a));
	a.ix = c->ix; a.paired = c->prm.paired; a.n_reads = n; a.roff = c->last_roff;
	a.cand_off = c->d_cand_off.as<int32_t>(); a.ncand = c->d_ncand.as<int32_t>(); a.cscore = c->d_cscore.as<int32_t>(); a.cpaired = c->d_cpaired.as<int32_t>();
	a.corient = c->d_corient.as<int32_t>(); a.cfrag = c->d_cfrag.as<int32_t>(); a.cnfrag = c->d_cnfrag.as<int32_t>(); a.rsum = c->d_rsum.as<ReadSum>();
	a.frags = c->d_frags.as<mc_frag_out>(); a.aln = c->d_aln.as<uint8_t>(); a.mapq_tab = d_tab.as<uint8_t>();
}

extern "C" {

int mc_map_batch(mc_ctx* c, const mc_batch_in* in, mc_batch_out* out)
{
	if (!c || !in || !out) { mc_set_error("mc_map_batch: null argument"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	// A large batch travels over PCIe in a few chunk-aligned pieces on the copy stream; the compute stream reverses mate 2 and
	// seeds piece i while piece i+1 is on the wire.  Everything after seeding sees the whole batch.
	const int pieces = in->n_reads >= 400000 && !getenv("MC_NO_PIECES") ? 4 : 1;
	ev_record(&c->ev[EV_START], c->stream);
	int rc = stage_reads(c, in, c->cur, pieces > 1 ? c->cstream : c->stream, pieces);
	if (rc) return rc;
	return run_batch(c, c->cur, out, true);
}

// ---- operator-level entries for the steps that only show up inside a mapped batch otherwise ------------------------------------
// Both run the pipeline of mc_map_batch on the caller's reads with the context's sequential state (totals, avgDist, SV-site
// bookkeeping, chunk grid, profile) set aside and put back afterwards: what comes out depends on the reads alone.
struct SavedState {
	mc_params prm; mc_totals tot; bool closed, discord_init; int64_t dg, dd; std::vector<mc_site_rec> inv, tnl;
	explicit SavedState(mc_ctx* c) : prm(c->prm), tot(c->tot), closed(c->library_closed), discord_init(c->discord_init), dg(c->discord_gpos), dd(c->discord_dist), inv(c->inv_sites), tnl(c->tnl_sites) {}
	void restore(mc_ctx* c) { c->prm = prm; c->tot = tot; c->library_closed = closed; c->discord_init = discord_init; c->discord_gpos = dg; c->discord_dist = dd; c->inv_sites = inv; c->tnl_sites = tnl; c->freeze_avg_dist = false; }
};

// SimplePairClustering -> RemoveRedundantAlnCan -> ProduceReadAlignment for n independent reads, i.e. the single-end branch of
// ReadMapping() (reference src/ReadMapping.cpp:575-585; ProduceReadAlignment src/ReadAlignment.cpp:306-430): candidates with their
// fragment lists, alignment strings and scores, AlnSummary per read.  Works on a paired context as well (mates are NOT reversed).
int mc_read_alignment_batch(mc_ctx* c, const mc_batch_in* in, mc_batch_out* out)
{
	if (!c || !in || !out) { mc_set_error("mc_read_alignment_batch: null argument"); return MC_ERR_ARG; }
	SavedState keep(c);
	c->prm.paired = 0; c->prm.update_profile = 0; c->prm.want_alignments = 1; c->prm.reserved[2] = 1; c->library_closed = false;
	memset(&c->tot, 0, sizeof(c->tot)); c->tot.avg_dist = 1000;
	const int rc = mc_map_batch(c, in, out);
	keep.restore(c);
	return rc;
}

// AlignmentRescue in its setting (reference src/AlignmentRescue.cpp:28-111, called from src/ReadMapping.cpp:466): every pair of the
// batch goes through IdentifySimplePairs, SimplePairClustering, the pairing / masking of ReadMapping(), AlignmentRescue(EstiDistance,
// read1, read2) and ProduceReadAlignment with EstiDistance = (int)(avg_dist * 1.5) for ALL pairs (src/ReadMapping.cpp:462) - no
// feedback from one 200-read chunk to the next, so a pair's records depend on the pair and on avg_dist alone.  Rescued candidates
// are the ones whose paired_idx points at a partner the FM-index seeds alone did not produce.
int mc_rescue_batch(mc_ctx* c, const mc_batch_in* in, uint32_t avg_dist, mc_batch_out* out)
{
	if (!c || !in || !out) { mc_set_error("mc_rescue_batch: null argument"); return MC_ERR_ARG; }
	if (!c->prm.paired) { mc_set_error("mc_rescue_batch: the context was created for single-end reads"); return MC_ERR_ARG; }
	SavedState keep(c);
	c->prm.update_profile = 0; c->prm.want_alignments = 1; c->prm.reserved[2] = 1; c->library_closed = false;
	memset(&c->tot, 0, sizeof(c->tot)); c->tot.avg_dist = avg_dist; c->freeze_avg_dist = true;
	const int rc = mc_map_batch(c, in, out);
	keep.restore(c);
	return rc;
}

// UpdateProfile / UpdateMultiHitCount as a step of its own (reference src/structure.h:248-249, src/AlignmentProfile.cpp:41-271;
// call site src/ReadMapping.cpp:559-568): with mc_defer_profile(ctx, 1) a mapped batch leaves its candidates, fragment lists and
// alignment strings in the device arenas and the profile untouched; mc_update_profile_last applies the reference's update to
// exactly those reads - dedup gate, strand / base / multi-hit counters, indel and break-point records - once.
int mc_defer_profile(mc_ctx* c, int32_t on)
{
	if (!c) { mc_set_error("mc_defer_profile: null context"); return MC_ERR_ARG; }
	c->defer_profile = on != 0;
	return MC_OK;
}
int mc_update_profile_last(mc_ctx* c)
{
	if (!c) { mc_set_error("mc_update_profile_last: null context"); return MC_ERR_ARG; }
	if (!c->prm.update_profile) { mc_set_error("mc_update_profile_last: context was created without update_profile"); return MC_ERR_ARG; }
	if (c->last_n <= 0 || c->last_profiled) { mc_set_error("mc_update_profile_last: no mapped batch on the device whose profile update is still due (mc_defer_profile, then mc_map_batch / mc_map_staged)"); return MC_ERR_ARG; }
	if (ordered_mode(c)) { mc_set_error("mc_update_profile_last: not with the ordered multi-GPU exchange (the dedup gate is exchanged inside mc_map_batch)"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	Ordered od;
	ev_record(&c->ev[EV_PROF0], c->stream);
	const int rc = profile_stage(c, c->last_a, c->last_n, od, c->last_frag_cap, c->last_hb);
	ev_record(&c->ev[EV_PROF1], c->stream);
	if (rc != MC_OK) return rc;
	if (dev_sync(c->stream)) return MC_ERR_CUDA;
	{ const double ms = ev_ms(&c->ev[EV_PROF0], &c->ev[EV_PROF1]); c->stats.ms_profile += ms; c->stats.ms_total += ms; }
	c->last_profiled = true;
	return MC_OK;
}

int mc_stage_batch_async(mc_ctx* c, const mc_batch_in* in, int32_t slot)
{
	if (!c || !in || slot < 0 || slot >= MC_SLOTS) { mc_set_error("mc_stage_batch: bad argument"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	Staged& st = c->slots[slot];
	int rc = stage_reads(c, in, st, c->cstream, 1);
	if (rc) return rc;
	// reverse-complement mate 2 once; the staged copy is then immutable
	PipeArgs a; memset(&a, 0, sizeof(a));
	a.pr.paired = c->prm.paired; a.seq = st.seq.as<uint8_t>() - st.base; a.roff = st.roff.as<int64_t>(); a.n_reads = in->n_reads;
	launch_prep(a, 0, in->n_reads, c->cstream);
#ifndef MC_HOSTEMU
	if (cuda_fail(cudaEventRecord(c->ev_slot[slot], c->cstream), "cudaEventRecord")) return MC_ERR_CUDA;
	st.pending = true;
#endif
	return MC_OK;
}

int mc_stage_batch(mc_ctx* c, const mc_batch_in* in, int32_t slot)
{
	int rc = mc_stage_batch_async(c, in, slot);
	if (rc) return rc;
	c->slots[slot].pending = false;
	return dev_sync(c->cstream) ? MC_ERR_CUDA : MC_OK;
}

int mc_map_staged(mc_ctx* c, int32_t slot, mc_batch_out* out)
{
	if (!c || !out || slot < 0 || slot >= MC_SLOTS || !c->slots[slot].valid) { mc_set_error("mc_map_staged: slot not staged"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
	if (c->slots[slot].pending)
	{
		if (cuda_fail(cudaStreamWaitEvent(c->stream, c->ev_slot[slot], 0), "cudaStreamWaitEvent")) return MC_ERR_CUDA;
		c->slots[slot].pending = false;
	}
#endif
	ev_record(&c->ev[EV_START], c->stream);
	return run_batch(c, c->slots[slot], out, false);
}

// Read ingest on the device: raw FASTQ text -> a staged batch (mc_map_staged maps it).  See include/mapcaller_b200.h.
int mc_ingest_fastq(mc_ctx* c, const mc_fastq_in* in, int32_t slot, mc_fastq_out* out)
{
	if (!c || !in || !out || slot < 0 || slot >= MC_SLOTS || !in->text1 || in->len1 < 0 || (in->text2 && in->len2 < 0)) { mc_set_error("mc_ingest_fastq: bad argument"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	const mc_stream_t s = c->cstream;   // the copy stream: another host thread may be mapping a different slot on c->stream
	memset(out, 0, sizeof(*out));
	Staged& st = c->slots[slot];
	st.valid = false; st.pending = false; st.n_pieces = 0; st.h_roff.clear(); st.has_text = false;
	const int nf = in->text2 ? 2 : 1;
	st.n_files = nf;
	const bool dbg = getenv("MC_DEBUG") != nullptr;
	const auto t_in = std::chrono::steady_clock::now();
	auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_in).count(); };
	double t_copy = 0, t_lines = 0;
#ifndef MC_HOSTEMU
	if (st.prefetched && cuda_fail(cudaStreamWaitEvent(s, c->ev_text[slot], 0), "cudaStreamWaitEvent")) return MC_ERR_CUDA;   // the copy queued by mc_ingest_prefetch
	if (dbg && st.prefetched) { cudaEventSynchronize(c->ev_text[slot]); t_copy = since(); }
#endif
	const uint8_t* text[2] = {in->text1, in->text2}; const int64_t len[2] = {in->len1, in->text2 ? in->len2 : 0};
	if (in->format != 0 && in->format != 1) { mc_set_error("mc_ingest_fastq: unknown format %d", in->format); return MC_ERR_ARG; }
	const int lpr = in->format == 1 ? 2 : 4;      // lines per record
	st.lpr = lpr;
	FastqArgs q; memset(&q, 0, sizeof(q)); q.lpr = lpr;
	DBuf* d_text = st.text; DBuf* d_cnt = c->fq_cnt; DBuf* d_off = c->fq_off; DBuf* d_lines = st.lines;
	DBuf& d_scan = c->fq_scan; DBuf& d_rlen = c->fq_rlen; DBuf& d_rsrc = c->fq_rsrc;   // kept between calls: no allocation in the steady state
	auto done = [&](int rc) { if (dbg) fprintf(stderr, "[mc] ingest slot %d: %.1f MB, text on the device after %.3f ms, lines counted after %.3f ms, staged after %.3f ms\n", slot, (in->len1 + (in->text2 ? in->len2 : 0)) / 1e6, t_copy, t_lines, since()); st.prefetched = false; return rc; };
	int bad = 0;
	int64_t n_tiles[2] = {0, 0}, n_lines[2] = {0, 0}; bool open_end[2] = {false, false};
	if (st.flag.reserve(16) || dev_zero(st.flag.p, 16, s)) return done(MC_ERR_CUDA);
	for (int f = 0; f < nf; f++)
	{
		n_tiles[f] = (len[f] + MC_FQ_TILE - 1) / MC_FQ_TILE;
		bad |= d_text[f].reserve(len[f] + 2 * MC_FQ_TILE) || d_cnt[f].reserve((n_tiles[f] + 1) * 4) || d_off[f].reserve((n_tiles[f] + 2) * 8) || d_scan.reserve(device_scan_scratch_bytes(n_tiles[f] + 1));
		if (bad) return done(MC_ERR_CUDA);
		if (!(st.prefetched && st.pre_src[f] == (const void*)text[f] && st.pre_len[f] == len[f])) bad |= upload(c, d_text[f].p, text[f], (size_t)len[f], s);
		q.text[f] = d_text[f].as<uint8_t>(); q.len[f] = len[f]; q.tile_cnt[f] = d_cnt[f].as<uint32_t>(); q.tile_off[f] = d_off[f].as<int64_t>();
		launch_fqcount(q, f, n_tiles[f], s);
		device_scan_u32(q.tile_cnt[f], d_off[f].as<int64_t>(), n_tiles[f], d_scan.as<int64_t>(), s);
		bad |= dev_d2h(&n_lines[f], d_off[f].as<int64_t>() + n_tiles[f], 8, s);
	}
	if (bad || dev_sync(s)) return done(MC_ERR_CUDA);
	t_lines = since();
	// whole records only (four lines each); at the end of the file a last line without newline still counts
	int64_t n_rec = -1;
	for (int f = 0; f < nf; f++)
	{
		open_end[f] = in->final_block && len[f] > 0 && text[f][len[f] - 1] != '\n';
		const int64_t rec = (n_lines[f] + (open_end[f] ? 1 : 0)) / lpr;
		(f == 0 ? out->records1 : out->records2) = rec;
		n_rec = n_rec < 0 ? rec : std::min(n_rec, rec);
	}
	int64_t n = nf == 2 ? 2 * n_rec : n_rec;
	if (in->max_reads > 0 && n > in->max_reads) n = in->max_reads;
	if (!in->final_block) n -= n % MC_CHUNK_READS;                      // keep the 200-read chunk grid of the reference intact
	if (c->prm.paired) n &= ~(int64_t)1;
	if (n >= (1ll << MC_KEY_SHIFT)) n = ((1ll << MC_KEY_SHIFT) - 1) / MC_CHUNK_READS * MC_CHUNK_READS;
	n_rec = nf == 2 ? n / 2 : n;
	out->n_reads = n;
	for (int f = 0; f < nf; f++)
	{
		bad |= d_lines[f].reserve((size_t)(n_lines[f] + 2) * 8);
		if (bad) return done(MC_ERR_CUDA);
		q.line_end[f] = d_lines[f].as<int64_t>() + 1; q.n_lines[f] = n_lines[f];
		const int64_t minus1 = -1;
		bad |= dev_h2d(d_lines[f].p, &minus1, 8, s);                      // line_end[-1] = -1: the first record starts at offset 0
		launch_fqlines(q, f, n_tiles[f], s);
		if (open_end[f]) bad |= dev_h2d(q.line_end[f] + n_lines[f], &len[f], 8, s);
	}
	if (n == 0)
	{
		if (dev_sync(s)) return done(MC_ERR_CUDA);
		st.n_reads = 0; st.n_bytes = 0; st.base = 0; st.n_slots = 0; st.valid = true;
		return done(MC_OK);
	}
	// where the records end: the caller carries the rest of the block over to the next call
	int64_t cons[2] = {0, 0};
	for (int f = 0; f < nf; f++)
	{
		const int64_t last_line = lpr * n_rec - 1;                         // index of the last consumed line
		if (last_line < n_lines[f]) bad |= dev_d2h(&cons[f], q.line_end[f] + last_line, 8, s); else cons[f] = len[f] - 1;
	}
	bad |= d_rlen.reserve((size_t)(n + 1) * 4) || d_rsrc.reserve((size_t)n * 8) || st.roff.reserve((size_t)(n + 1) * 8) || d_scan.reserve(device_scan_scratch_bytes(n + 1));
	if (bad) return done(MC_ERR_CUDA);
	q.n_reads = n; q.rlen = d_rlen.as<uint32_t>(); q.rsrc = d_rsrc.as<int64_t>(); q.flag = st.flag.as<mc_u64>();
	launch_fqread(q, n, s);
	device_scan_u32(q.rlen, st.roff.as<int64_t>(), n, d_scan.as<int64_t>(), s);
	int64_t n_bases = 0;
	bad |= dev_d2h(&n_bases, st.roff.as<int64_t>() + n, 8, s);
	if (bad || dev_sync(s)) return done(MC_ERR_CUDA);
	out->consumed1 = cons[0] + 1; out->consumed2 = nf == 2 ? cons[1] + 1 : 0; out->n_bases = n_bases;
	st.n_reads = n; st.n_bytes = n_bases; st.base = 0; st.n_slots = n_bases / 17 + n;
	bad |= st.seq.reserve((size_t)n_bases + 16) || st.seed_off.reserve((size_t)(n + 2) * 8) || st.cap.reserve((size_t)n * 4) || st.scan.reserve(device_scan_scratch_bytes(n));
	if (bad) return done(MC_ERR_CUDA);
	q.roff = st.roff.as<int64_t>(); q.seq = st.seq.as<uint8_t>();
	launch_fqcopy(q, n, s);
	CapArgs cq; cq.roff = st.roff.as<int64_t>(); cq.cap = st.cap.as<uint32_t>(); cq.flag = st.flag.as<mc_u64>();
	launch_seedcap(cq, n, s);
	device_scan_u32(cq.cap, st.seed_off.as<int64_t>(), n, st.scan.as<int64_t>(), s);
	PipeArgs a; memset(&a, 0, sizeof(a));
	a.pr.paired = c->prm.paired; a.seq = st.seq.as<uint8_t>(); a.roff = st.roff.as<int64_t>(); a.n_reads = n;
	launch_prep(a, 0, n, s);
	if (c->prm.want_alignments) { st.h_roff.resize((size_t)n + 1); bad |= dev_d2h(st.h_roff.data(), st.roff.p, (size_t)(n + 1) * 8, s); }
	mc_u64 flag = 0;
	bad |= dev_d2h(&flag, st.flag.p, 8, s);
	if (bad || dev_sync(s)) return done(MC_ERR_CUDA);
	if (flag >> 57) { mc_set_error("mc_ingest_fastq: FASTA records with the sequence on several lines (or text that is not FASTA): only one line of bases per record is parsed on the device"); return done(MC_ERR_ARG); }
	if (flag) { mc_set_error("mc_ingest_fastq: a record has an empty read or one longer than %d bases", MC_MAX_RLEN); return done(MC_ERR_ARG); }
	st.valid = true; st.has_text = true;
	return done(MC_OK);
}

// Queues the host -> device copy of the FASTQ blocks the next mc_ingest_fastq(slot) is going to parse and returns: plain DMA
// on a stream of its own, so the PCIe link stays busy while an earlier block is parsed and a still earlier one is mapped.
int mc_ingest_prefetch(mc_ctx* c, const mc_fastq_in* in, int32_t slot)
{
	if (!c || !in || slot < 0 || slot >= MC_SLOTS || !in->text1 || in->len1 < 0 || (in->text2 && in->len2 < 0)) { mc_set_error("mc_ingest_prefetch: bad argument"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	Staged& st = c->slots[slot];
	st.valid = false; st.has_text = false; st.prefetched = false;
	const uint8_t* text[2] = {in->text1, in->text2}; const int64_t len[2] = {in->len1, in->text2 ? in->len2 : 0};
	for (int f = 0; f < (in->text2 ? 2 : 1); f++)
	{
		if (st.text[f].reserve((size_t)len[f] + 2 * MC_FQ_TILE) || upload(c, st.text[f].p, text[f], (size_t)len[f], c->dstream)) return MC_ERR_CUDA;
		st.pre_src[f] = text[f]; st.pre_len[f] = len[f];
	}
#ifndef MC_HOSTEMU
	if (cuda_fail(cudaEventRecord(c->ev_text[slot], c->dstream), "cudaEventRecord")) return MC_ERR_CUDA;
#endif
	st.prefetched = true;
	return MC_OK;
}

// SAM text of the batch just mapped from `slot` (reference src/SamReport.cpp:324-488), assembled on the device from the
// batch's arenas and the FASTQ text mc_ingest_fastq left in the slot.  See include/mapcaller_b200.h.
int mc_sam_text(mc_ctx* c, int32_t slot, int32_t all_best, const uint8_t** text, int64_t* n_bytes)
{
	if (!c || !text || !n_bytes || slot < 0 || slot >= MC_SLOTS) { mc_set_error("mc_sam_text: bad argument"); return MC_ERR_ARG; }
	Staged& st = c->slots[slot];
	if (!st.has_text || c->last_n <= 0 || c->last_roff != st.roff.as<int64_t>()) { mc_set_error("mc_sam_text: slot %d does not hold the FASTQ text of the batch just mapped (mc_ingest_fastq + mc_map_staged)", slot); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	const int64_t n = c->last_n; mc_stream_t s = c->stream;
	DBuf &d_tab = c->d_sam[0], &d_len = c->d_sam[2], &d_off = c->d_sam[3], &d_txt = c->d_sam[4];
	if (sam_mapq_table(c, d_tab, s)) return MC_ERR_CUDA;
	if (c->d_chrom_names.cap == 0)
	{
		// chromosome names for the RNAME column: back to back, n_chrom + 1 offsets
		std::vector<uint8_t> names; std::vector<int32_t> off(1, 0);
		for (const std::string& nm : c->chrom_names) { names.insert(names.end(), nm.begin(), nm.end()); off.push_back((int32_t)names.size()); }
		if (c->d_chrom_names.reserve(names.size() + 16) || c->d_chrom_name_off.reserve(off.size() * 4) || dev_h2d(c->d_chrom_names.p, names.data(), names.size(), s) || dev_h2d(c->d_chrom_name_off.p, off.data(), off.size() * 4, s) || dev_sync(s)) return MC_ERR_CUDA;
	}
	if (d_len.reserve((size_t)(n + 1) * 4) || d_off.reserve((size_t)(n + 2) * 8) || c->d_scan.reserve(device_scan_scratch_bytes(n))) return MC_ERR_CUDA;
	SamTextArgs t; memset(&t, 0, sizeof(t));
	sam_args_of(c, t.s, n, d_tab);
	for (int f = 0; f < st.n_files; f++) { t.text[f] = st.text[f].as<uint8_t>(); t.line_end[f] = st.lines[f].as<int64_t>() + 1; }
	t.two_files = st.n_files == 2; t.unique = all_best ? 0 : 1; t.lpr = st.lpr;
	t.chrom_names = c->d_chrom_names.as<uint8_t>(); t.chrom_name_off = c->d_chrom_name_off.as<int32_t>();
	t.tlen = d_len.as<uint32_t>(); t.toff = d_off.as<int64_t>();
	launch_samtext(t, n, false, s);
	device_scan_u32(t.tlen, d_off.as<int64_t>(), n, c->d_scan.as<int64_t>(), s);
	int64_t total = 0;
	if (dev_d2h(&total, d_off.as<int64_t>() + n, 8, s) || dev_sync(s)) return MC_ERR_CUDA;
	if (d_txt.reserve((size_t)total + 16) || c->h_sam_text.reserve((size_t)total + 16)) return MC_ERR_CUDA;
	t.out = d_txt.as<uint8_t>();
	launch_samtext(t, n, true, s);
	if (dev_d2h(c->h_sam_text.p, d_txt.p, (size_t)total, s) || dev_sync(s)) return MC_ERR_CUDA;
	*text = c->h_sam_text.as<uint8_t>(); *n_bytes = total;
	return MC_OK;
}

// Block totals of the six difference-array lanes over the OWNED columns and their exclusive prefixes (pre[k * nb + b]).  After a
// reduce-scatter the owned range starts in the middle of the genome: the sums of everything before it (own_carry) are planted
// in the block just before the range and the scan starts there.
static int profile_prefix(mc_ctx* c, const DevProfile& p, int64_t nb, int64_t* sums, mc_stream_t s)
{
	const int64_t ob0 = c->own_beg / MC_PROF_BLOCK, ob1 = (c->own_end + MC_PROF_BLOCK - 1) / MC_PROF_BLOCK;
	launch_profsum(p, c->G, nb, ob0, ob1, sums, s);
	const int64_t s0 = ob0 > 0 ? ob0 - 1 : 0;
	int bad = 0;
	if (ob0 > 0) for (int k = 0; k < 6; k++) bad |= dev_h2d(sums + k * nb + s0, &c->own_carry[k], 8, s);
	for (int k = 0; k < 6; k++) device_exscan_i64(sums + k * nb + s0, ob1 - s0, sums + 6 * nb + k, c->d_scan2.p, s);
	return bad;
}

// packs the columns [beg, end) tile by tile; every tile is either copied to `outp` or reduced into the four counters `acc`
static int profile_walk(mc_ctx* c, int64_t beg, int64_t end, void* outp, mc_u64* d_acc, bool checksum = false)
{
	DevProfile p; p.base16 = c->d_base16.as<uint32_t>(); p.sdiff = c->d_sdiff.as<int32_t>(); p.cdiff = c->d_cdiff.as<int32_t>(); p.mdiff = c->d_mdiff.as<int32_t>();
	p.rcount = c->d_rcount.as<uint8_t>();
	// the range counters are difference arrays: block totals, exclusive scan over the blocks, then every block sums its own columns
	const int64_t nb = (c->G + MC_PROF_BLOCK - 1) / MC_PROF_BLOCK;
	DBuf d_sums;
	if (d_sums.reserve((size_t)(6 * nb + 8) * 8) || c->d_scan2.reserve(device_scan_scratch_bytes(nb))) return MC_ERR_CUDA;
	int64_t* sums = d_sums.as<int64_t>();
	if (beg < c->own_beg) beg = c->own_beg;            // (mc_profile_read refuses such ranges; the reductions cover what this rank owns)
	if (end > c->own_end) end = c->own_end;
	if (profile_prefix(c, p, nb, sums, c->stream)) { d_sums.release(); return MC_ERR_CUDA; }
	const int64_t tile_blocks = (1 << 24) / MC_PROF_BLOCK;
	int rc = MC_OK;
	if (c->d_sort.reserve((size_t)(tile_blocks * MC_PROF_BLOCK) * 16)) rc = MC_ERR_CUDA;
	for (int64_t b0 = beg / MC_PROF_BLOCK; rc == MC_OK && b0 * MC_PROF_BLOCK < end; b0 += tile_blocks)
	{
		const int64_t b1 = std::min(nb, b0 + tile_blocks);
		const int64_t tb = std::max(beg, b0 * MC_PROF_BLOCK), te = std::min(end, b1 * MC_PROF_BLOCK);
		launch_profpack(c->ix, p, nb, sums, b0, b1, tb, te, c->d_sort.as<uint64_t>(), c->stream);
		if (outp) { if (dev_d2h((uint8_t*)outp + (tb - beg) * 16, c->d_sort.p, (size_t)(te - tb) * 16, c->stream) || dev_sync(c->stream)) rc = MC_ERR_CUDA; }
		else if (checksum) launch_profhash(tb, te - tb, c->d_sort.as<uint64_t>(), d_acc, c->stream);
		else launch_profstat(te - tb, c->d_sort.as<uint64_t>(), d_acc, c->stream);
	}
	if (dev_sync(c->stream)) rc = MC_ERR_CUDA;
	d_sums.release();
	return rc;
}

int mc_profile_read(mc_ctx* c, int64_t beg, int64_t end, void* outp)
{
	if (!c || !outp || beg < 0 || end > c->G || beg > end) { mc_set_error("mc_profile_read: bad range"); return MC_ERR_ARG; }
	if (!c->prm.update_profile) { mc_set_error("mc_profile_read: context was created without update_profile"); return MC_ERR_ARG; }
	if (beg == end) return MC_OK;
	if (beg < c->own_beg || end > c->own_end) { mc_set_error("mc_profile_read: after mc_profile_reduce_scatter this rank holds the columns [%lld, %lld) only", (long long)c->own_beg, (long long)c->own_end); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	return profile_walk(c, beg, end, outp, nullptr);
}

int mc_profile_summary(mc_ctx* c, mc_profile_stats* out)
{
	if (!c || !out) { mc_set_error("mc_profile_summary: null argument"); return MC_ERR_ARG; }
	if (!c->prm.update_profile) { mc_set_error("mc_profile_summary: context was created without update_profile"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	DBuf d_acc;
	if (d_acc.reserve(64) || dev_zero(d_acc.p, 64, c->stream)) return MC_ERR_CUDA;
	int rc = profile_walk(c, 0, c->G, nullptr, d_acc.as<mc_u64>());
	mc_u64 h[4] = {0, 0, 0, 0};
	if (rc == MC_OK && (dev_d2h(h, d_acc.p, 32, c->stream) || dev_sync(c->stream))) rc = MC_ERR_CUDA;
	d_acc.release();
#ifndef MC_HOSTEMU
	if (rc == MC_OK && c->scattered)    // every rank reduced the tile it owns: the sums of all ranks are the library's
	{
		std::vector<mc_u64> all;
		if (gather_words(c, h, 4, all)) return MC_ERR_NCCL;
		for (int k = 0; k < 4; k++) { h[k] = 0; for (int r = 0; r < c->comm_size; r++) h[k] += all[(size_t)4 * r + k]; }
	}
#endif
	out->aligned_bases = (int64_t)h[0]; out->coverage_sum = (int64_t)h[1]; out->dup_sites = (int64_t)h[2]; out->dup_reads = (int64_t)h[3];
	return rc;
}

int mc_profile_checksum(mc_ctx* c, uint64_t out[2])
{
	if (!c || !out) { mc_set_error("mc_profile_checksum: null argument"); return MC_ERR_ARG; }
	if (!c->prm.update_profile) { mc_set_error("mc_profile_checksum: context was created without update_profile"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	DBuf d_acc;
	if (d_acc.reserve(64) || dev_zero(d_acc.p, 64, c->stream)) return MC_ERR_CUDA;
	int rc = profile_walk(c, 0, c->G, nullptr, d_acc.as<mc_u64>(), true);
	mc_u64 h[2] = {0, 0};
	if (rc == MC_OK && (dev_d2h(h, d_acc.p, 16, c->stream) || dev_sync(c->stream))) rc = MC_ERR_CUDA;
	d_acc.release();
#ifndef MC_HOSTEMU
	if (rc == MC_OK && c->scattered)    // sum and xor of the per-column hashes combine over the tiles
	{
		std::vector<mc_u64> all;
		if (gather_words(c, h, 2, all)) return MC_ERR_NCCL;
		h[0] = h[1] = 0;
		for (int r = 0; r < c->comm_size; r++) { h[0] += all[(size_t)2 * r]; h[1] ^= all[(size_t)2 * r + 1]; }
	}
#endif
	out[0] = h[0]; out[1] = h[1];
	return rc;
}

int mc_profile_indels(mc_ctx* c, const mc_indel_rec** recs, int64_t* n_recs, const uint8_t** seq_arena)
{
#ifndef MC_HOSTEMU
	if (c) cudaSetDevice(c->prm.device);   // a process may hold contexts on several GPUs
#endif
	if (!c || !recs || !n_recs || !seq_arena) { mc_set_error("mc_profile_indels: null argument"); return MC_ERR_ARG; }
	PersistBumps pb;
	mc_stream_t s = c->stream;
	if (dev_d2h(&pb, c->d_pbump.p, sizeof(pb), s) || dev_sync(s)) return MC_ERR_CUDA;
	const int64_t n = (int64_t)pb.ind;
	c->ind_out.clear(); c->ind_seq_out.clear();
	if (n >= 0xffffffffll) { mc_set_error("mc_profile_indels: more than 2^32 raw records"); return MC_ERR_OVERFLOW; }
	if (n > 0)
	{
		// the device sorts the raw records by (kind, position, hash of the sequence) and lays them and their sequences out in that
		// order; what is left for the host is one sequential pass
		DBuf &d_key = c->d_ia[0], &d_ktmp = c->d_ia[1], &d_idx = c->d_ia[2], &d_itmp = c->d_ia[3], &d_len = c->d_ia[4], &d_off = c->d_ia[5], &d_rec = c->d_ia[6], &d_seq = c->d_ia[7], &d_scr = c->d_ia[8];
		int bad = d_key.reserve((size_t)n * 8) || d_ktmp.reserve((size_t)n * 8) || d_idx.reserve((size_t)n * 4) || d_itmp.reserve((size_t)n * 4) || d_len.reserve((size_t)(n + 1) * 4) || d_off.reserve((size_t)(n + 2) * 8);
		bad = bad || d_rec.reserve((size_t)n * sizeof(mc_indel_rec)) || d_seq.reserve((size_t)pb.ind_seq + 16) || d_scr.reserve(device_sort_pairs_scratch_bytes(n)) || c->d_scan2.reserve(device_scan_scratch_bytes(n));
		if (bad) return MC_ERR_CUDA;
		launch_indkey(n, c->d_ind.as<mc_indel_rec>(), c->d_ind_seq.as<uint8_t>(), d_key.as<uint64_t>(), d_idx.as<uint32_t>(), s);
		device_sort_pairs(d_key.as<uint64_t>(), d_ktmp.as<uint64_t>(), d_idx.as<uint32_t>(), d_itmp.as<uint32_t>(), n, d_scr.p, d_scr.cap, s);
		launch_indlen(n, c->d_ind.as<mc_indel_rec>(), d_idx.as<uint32_t>(), d_len.as<uint32_t>(), s);
		device_scan_u32(d_len.as<uint32_t>(), d_off.as<int64_t>(), n, c->d_scan2.as<int64_t>(), s);
		launch_indgather(n, c->d_ind.as<mc_indel_rec>(), c->d_ind_seq.as<uint8_t>(), d_idx.as<uint32_t>(), d_off.as<int64_t>(), d_rec.as<mc_indel_rec>(), d_seq.as<uint8_t>(), s);
		if (c->h_ind_rec.reserve((size_t)n * sizeof(mc_indel_rec)) || c->h_ind_seq.reserve((size_t)pb.ind_seq + 16)) return MC_ERR_CUDA;
		if (dev_d2h(c->h_ind_rec.p, d_rec.p, (size_t)n * sizeof(mc_indel_rec), s) || dev_d2h(c->h_ind_seq.p, d_seq.p, (size_t)pb.ind_seq, s) || dev_sync(s)) return MC_ERR_CUDA;
		const mc_indel_rec* raw = c->h_ind_rec.as<mc_indel_rec>(); const uint8_t* sq = c->h_ind_seq.as<uint8_t>();
		// aggregate like map<int64, map<string, uint16_t>>::operator[]++ (reference src/AlignmentProfile.cpp:123,129): the records of
		// one (kind, position) are adjacent; their distinct sequences come out in the order std::string compares them
		auto seq_less = [&](const mc_indel_rec* a, const mc_indel_rec* b) {
			const int m = memcmp(sq + a->seq_off, sq + b->seq_off, (size_t)std::min(a->len, b->len));
			return m != 0 ? m < 0 : a->len < b->len;
		};
		std::vector<const mc_indel_rec*> grp;
		c->ind_out.reserve((size_t)n / 4 + 16); c->ind_seq_out.reserve((size_t)pb.ind_seq / 4 + 16);
		for (int64_t i = 0; i < n;)
		{
			int64_t j = i + 1;
			while (j < n && raw[j].kind == raw[i].kind && raw[j].pos == raw[i].pos) j++;
			grp.clear();
			for (int64_t k = i; k < j; k++) grp.push_back(raw + k);
			// equal sequences are already adjacent (same hash) except after a hash collision: a sort of the group's run heads would
			// do, sorting the whole (small) group is simpler and exact either way
			if (j - i > 1) std::sort(grp.begin(), grp.end(), seq_less);
			for (size_t g = 0; g < grp.size();)
			{
				size_t h = g + 1;
				while (h < grp.size() && !seq_less(grp[g], grp[h])) h++;
				mc_indel_rec r; r.pos = grp[g]->pos; r.kind = grp[g]->kind; r.len = grp[g]->len; r.count = (int32_t)((h - g) & 0xFFFF);
				r.seq_off = (int32_t)c->ind_seq_out.size();
				c->ind_seq_out.insert(c->ind_seq_out.end(), sq + grp[g]->seq_off, sq + grp[g]->seq_off + grp[g]->len);
				c->ind_out.push_back(r);
				g = h;
			}
			i = j;
		}
	}
	c->ind_seq_out.push_back(0);
	*recs = c->ind_out.data(); *n_recs = (int64_t)c->ind_out.size(); *seq_arena = c->ind_seq_out.data();
	return MC_OK;
}

int mc_profile_breakpoints(mc_ctx* c, const mc_breakpoint_rec** recs, int64_t* n_recs)
{
#ifndef MC_HOSTEMU
	if (c) cudaSetDevice(c->prm.device);   // a process may hold contexts on several GPUs
#endif
	if (!c || !recs || !n_recs) { mc_set_error("mc_profile_breakpoints: null argument"); return MC_ERR_ARG; }
	PersistBumps pb;
	if (dev_d2h(&pb, c->d_pbump.p, sizeof(pb), c->stream) || dev_sync(c->stream)) return MC_ERR_CUDA;
	std::vector<int64_t> raw((size_t)pb.bp);
	if (dev_d2h(raw.data(), c->d_bp.p, raw.size() * 8, c->stream) || dev_sync(c->stream)) return MC_ERR_CUDA;
	std::map<int64_t, uint32_t> agg;
	for (size_t i = 0; i < raw.size(); i++) agg[raw[i]]++;
	c->bp_out.clear();
	for (auto& kv : agg) c->bp_out.push_back({kv.first, (int64_t)(kv.second & 0xFFFF)});
	*recs = c->bp_out.data(); *n_recs = (int64_t)c->bp_out.size();
	return MC_OK;
}

// Records in the order the reference's thread body pushes them (src/ReadMapping.cpp:493,502,513-519); the
// caller applies the thread-end std::sort by gPos (:629-630) itself so that ties keep the reference's order.
int mc_profile_sites(mc_ctx* c, int32_t kind, const mc_site_rec** recs, int64_t* n_recs)
{
#ifndef MC_HOSTEMU
	if (c) cudaSetDevice(c->prm.device);   // a process may hold contexts on several GPUs
#endif
	if (!c || !recs || !n_recs || kind < 0 || kind > 1) { mc_set_error("mc_profile_sites: bad argument"); return MC_ERR_ARG; }
	std::vector<mc_site_rec>& v = kind == 0 ? c->inv_sites : c->tnl_sites;
	*recs = v.data(); *n_recs = (int64_t)v.size();
	return MC_OK;
}

// ---- SAM record fields of the last batch (reference src/SamReport.cpp:7-316, 324-488) -------------------------------
int mc_sam_records(mc_ctx* c, const mc_sam_rec** recs, int64_t* n_recs, const uint8_t** cigar_arena)
{
	if (!c || !recs || !n_recs || !cigar_arena) { mc_set_error("mc_sam_records: null argument"); return MC_ERR_ARG; }
	if (c->last_n <= 0) { mc_set_error("mc_sam_records: no mapped batch on the device (call it right after mc_map_batch / mc_map_staged)"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	const int64_t n = c->last_n; mc_stream_t s = c->stream;
	DBuf &d_tab = c->d_sam[0], &d_out = c->d_sam[1], &d_len = c->d_sam[2], &d_off = c->d_sam[3], &d_cig = c->d_sam[4];
	int bad = sam_mapq_table(c, d_tab, s) || d_out.reserve((size_t)n * sizeof(mc_sam_rec)) || d_len.reserve((size_t)(n + 1) * 4) || d_off.reserve((size_t)(n + 2) * 8) || c->d_scan.reserve(device_scan_scratch_bytes(n));
	if (bad) return MC_ERR_CUDA;
	SamArgs a; sam_args_of(c, a, n, d_tab);
	a.out = d_out.as<mc_sam_rec>(); a.clen = d_len.as<uint32_t>(); a.coff = d_off.as<int64_t>();
	launch_samrec(a, n, false, s);
	device_scan_u32(a.clen, d_off.as<int64_t>(), n, c->d_scan.as<int64_t>(), s);
	int64_t total = 0;
	if (dev_d2h(&total, d_off.as<int64_t>() + n, 8, s) || dev_sync(s)) return MC_ERR_CUDA;
	if (total > 0x7fffffffll) { mc_set_error("mc_sam_records: CIGAR text of the batch exceeds 2 GiB"); return MC_ERR_OVERFLOW; }
	if (d_cig.reserve((size_t)total + 16)) return MC_ERR_CUDA;
	a.cigar = d_cig.as<uint8_t>();
	launch_samrec(a, n, true, s);
	c->sam_out.resize((size_t)n); c->sam_cigar.resize((size_t)total + 1);
	if (dev_d2h(c->sam_out.data(), d_out.p, (size_t)n * sizeof(mc_sam_rec), s) || dev_d2h(c->sam_cigar.data(), d_cig.p, (size_t)total, s) || dev_sync(s)) return MC_ERR_CUDA;
	c->sam_cigar[(size_t)total] = 0;
	*recs = c->sam_out.data(); *n_recs = n; *cigar_arena = c->sam_cigar.data();
	return MC_OK;
}

// ---- variant-calling scan (reference src/VariantCalling.cpp:106-120, 550-680) ---------------------------------
void mc_vc_params_default(mc_vc_params* p)
{
	memset(p, 0, sizeof(*p));
	p->min_allele_depth = 5; p->frequency_thr = 0.2f; p->ploidy = 2; p->min_cnv_size = 50; p->min_unmapped_size = 50;
}

// GetAreaIndFrequency (src/VariantCalling.cpp:60-98) for every key of one indel map: recs[i0, i1) are the aggregated
// records of one kind in (pos, sequence) order, i.e. the iteration order of map<int64_t, map<string, uint16_t>>.
// A key is a candidate when the window's winner sits on it (the function returns 0 everywhere else).
static void vc_window_join(const std::vector<mc_indel_rec>& recs, size_t i0, size_t i1, std::vector<VcCand>& out)
{
	size_t lo = i0, hi = i0;
	for (size_t i = i0; i < i1; i++)
	{
		const int64_t g = recs[i].pos;
		if (i > i0 && recs[i - 1].pos == g) continue;
		while (lo < i1 && recs[lo].pos < g - 5) lo++;
		while (hi < i1 && recs[hi].pos <= g + 5) hi++;
		int64_t max_pos = 0; int freq = 0, max_freq = 0, best_len = 0, best_off = 0;
		for (size_t k = lo; k < hi; k++)
		{
			const int cnt = recs[k].count;
			freq += cnt;
			if (max_freq < cnt) { max_freq = cnt; best_len = recs[k].len; best_off = recs[k].seq_off; max_pos = recs[k].pos; }
			else if (max_freq == cnt && recs[k].len > best_len) { best_len = recs[k].len; best_off = recs[k].seq_off; max_pos = recs[k].pos; }
		}
		if (max_pos == g) { VcCand c; c.pos = g; c.kind = recs[i].kind; c.freq = freq; c.alt_off = best_off; c.alt_len = best_len; out.push_back(c); }
	}
}

int mc_variant_scan(mc_ctx* c, const mc_vc_params* vp, const mc_variant_rec** recs, int64_t* n_recs, const uint8_t** alt_arena, const int32_t** block_depth, int64_t* n_blocks)
{
	if (!c || !vp || !recs || !n_recs || !alt_arena || !block_depth || !n_blocks) { mc_set_error("mc_variant_scan: null argument"); return MC_ERR_ARG; }
	if (!c->prm.update_profile) { mc_set_error("mc_variant_scan: context was created without update_profile"); return MC_ERR_ARG; }
	if (vp->min_allele_depth < 1) { mc_set_error("mc_variant_scan: min_allele_depth must be >= 1"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	// indel window winners from the aggregated records (sparse; sorted by (kind, pos, sequence))
	const mc_indel_rec* ir; int64_t nir; const uint8_t* arena;
	int rc = mc_profile_indels(c, &ir, &nir, &arena);
	if (rc != MC_OK) return rc;
	std::vector<VcCand> cand;
	{
		size_t split = 0; while (split < c->ind_out.size() && c->ind_out[split].kind == 0) split++;
		vc_window_join(c->ind_out, 0, split, cand); vc_window_join(c->ind_out, split, c->ind_out.size(), cand);
		std::sort(cand.begin(), cand.end(), [](const VcCand& x, const VcCand& y) { return x.pos != y.pos ? x.pos < y.pos : x.kind < y.kind; });
	}
	mc_stream_t s = c->stream;
	DevProfile p; p.base16 = c->d_base16.as<uint32_t>(); p.sdiff = c->d_sdiff.as<int32_t>(); p.cdiff = c->d_cdiff.as<int32_t>(); p.mdiff = c->d_mdiff.as<int32_t>();
	p.rcount = c->d_rcount.as<uint8_t>();
	const int64_t G = c->G, nb = (G + MC_PROF_BLOCK - 1) / MC_PROF_BLOCK, nvb = (G + MC_VC_BLOCK - 1) / MC_VC_BLOCK;
	// After mc_profile_reduce_scatter this rank scans the columns it owns, [R0, R1) (R0 a multiple of MC_TILE_ALIGN); the run
	// carriers of the ranks before it come in through one small all-gather and the records of all ranks are gathered at the
	// end, so every rank returns the scan of the whole genome.  Otherwise [R0, R1) = [0, G).
	const int64_t R0 = c->own_beg, R1 = c->own_end;
	const int64_t vbF = R0 / MC_VC_BLOCK, vbL = (R1 + MC_VC_BLOCK - 1) / MC_VC_BLOCK, nvo = vbL - vbF;   // blocks of 100 columns of the range
	bool dist = false;
#ifndef MC_HOSTEMU
	dist = c->scattered;
#endif
	if (dist && vp->gvcf && !vp->monomorphic) { mc_set_error("mc_variant_scan: a gVCF scan needs the whole profile on one rank (mc_profile_allreduce instead of mc_profile_reduce_scatter)"); return MC_ERR_ARG; }
	// The scan makes three passes over the packed profile (depths, counts, records).  When HBM has room for the whole
	// MappingRecord_t image (16 bytes per column) it is packed once and kept; otherwise it is re-packed tile by tile per pass.
	int64_t tile_cols = (int64_t)25600 * 640;           // a multiple of MC_PROF_BLOCK and of MC_VC_BLOCK
#ifndef MC_HOSTEMU
	{
		size_t free_b = 0, total_b = 0;
		const size_t whole = (size_t)((R1 - R0 + 25599) / 25600 + 1) * 25600 * 16;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (c->d_sort.cap >= whole || free_b > whole + ((size_t)8 << 30)) && !getenv("MC_VC_TILED")) tile_cols = (int64_t)(whole / 16);
	}
#endif
	const int64_t n_tiles = (R1 - R0 + tile_cols - 1) / tile_cols;
	DBuf &d_sums = c->d_vc[0], &d_depth = c->d_vc[1], &d_ng = c->d_vc[2], &d_nd = c->d_vc[3], &d_ln = c->d_vc[4], &d_le = c->d_vc[5], &d_lead = c->d_vc[6],
	     &d_cnt = c->d_vc[7], &d_off = c->d_vc[8], &d_scan = c->d_vc[9], &d_cand = c->d_vc[10], &d_out = c->d_vc[11];
	const bool trace = getenv("MC_VC_TRACE") != nullptr;   // stage timings on stderr (each stage ends with a stream sync then)
	auto t_last = std::chrono::steady_clock::now();
	auto mark = [&](const char* what) {
		if (!trace) return;
		dev_sync(s);
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[mc_variant_scan] %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
		t_last = now;
	};
	auto done = [&](int r) { dev_sync(s); return r; };
	mark("indels");
	const int64_t nvb_pad = nvb + MC_TILE_PAD / MC_VC_BLOCK;      // room for the equal per-rank chunks of the depth all-gather
	int bad = d_sums.reserve((size_t)(6 * nb + 8) * 8) || d_depth.reserve((size_t)nvb_pad * 4) || d_ng.reserve((size_t)nvb * 8) || d_nd.reserve((size_t)nvb * 8);
	bad |= d_ln.reserve((size_t)nvb * 8) || d_le.reserve((size_t)nvb * 8) || d_lead.reserve((size_t)nvb * 4) || d_cnt.reserve((size_t)(nvb + 1) * 4);
	bad |= d_off.reserve((size_t)(nvb + 2) * 8) || d_scan.reserve(device_scan_scratch_bytes(nvb)) || d_cand.reserve((cand.size() + 1) * sizeof(VcCand));
	bad |= c->d_sort.reserve((size_t)std::min(tile_cols, nb * MC_PROF_BLOCK) * 16) || c->d_scan2.reserve(device_scan_scratch_bytes(std::max(nb, nvb)));
	if (bad) return done(MC_ERR_CUDA);
	if (dev_h2d(d_cand.p, cand.data(), cand.size() * sizeof(VcCand), s)) return done(MC_ERR_CUDA);
	int64_t* sums = d_sums.as<int64_t>();
	if (profile_prefix(c, p, nb, sums, s)) return done(MC_ERR_CUDA);
	VcArgs a; memset(&a, 0, sizeof(a));
	a.vp = *vp; if (a.vp.gvcf && a.vp.monomorphic) a.vp.gvcf = 0;   // src/main.cpp:322
	a.ix = c->ix; a.G = G; a.n_blocks = nvb; a.recs = c->d_sort.as<uint64_t>();
	a.depth = d_depth.as<int32_t>(); a.last_nongap = d_ng.as<int64_t>(); a.last_nondup = d_nd.as<int64_t>(); a.last_normal = d_ln.as<int64_t>(); a.last_event = d_le.as<int64_t>();
	a.lead_min = d_lead.as<int32_t>(); a.cnt = d_cnt.as<uint32_t>(); a.off = d_off.as<int64_t>(); a.cand = d_cand.as<VcCand>(); a.n_cand = (int64_t)cand.size();
	int64_t packed = -1;
	auto pack = [&](int64_t t) {   // MappingRecord_t image of tile t in d_sort (kept when the range is a single tile)
		a.tile_beg = R0 + t * tile_cols; a.tile_end = std::min(R1, a.tile_beg + tile_cols);
		if (packed == t) return;
		launch_profpack(c->ix, p, nb, sums, a.tile_beg / MC_PROF_BLOCK, (a.tile_end + MC_PROF_BLOCK - 1) / MC_PROF_BLOCK, a.tile_beg, a.tile_end, c->d_sort.as<uint64_t>(), s);
		packed = t;
	};
	auto vb0 = [&](int64_t t) { return vbF + t * (tile_cols / MC_VC_BLOCK); };
	auto vb1 = [&](int64_t t) { return std::min(vbL, vbF + (t + 1) * (tile_cols / MC_VC_BLOCK)); };
	mark("prefix");
	for (int64_t t = 0; t < n_tiles; t++) { pack(t); launch_vcdepth(a, vb0(t), vb1(t), s); }
	mark("pack+depth");
	if (nvo > 0) { device_incmax_i64(a.last_nongap + vbF, nvo, c->d_scan2.p, s); device_incmax_i64(a.last_nondup + vbF, nvo, c->d_scan2.p, s); }
#ifndef MC_HOSTEMU
	if (dist)
	{
		// run carriers: the last non-gap / non-dup column of everything before this rank's range = the maximum over the earlier ranks
		mc_u64 mine[2] = {(mc_u64)-1, (mc_u64)-1};       // -1 (none) as int64
		if (nvo > 0 && (dev_d2h(&mine[0], a.last_nongap + vbL - 1, 8, s) || dev_d2h(&mine[1], a.last_nondup + vbL - 1, 8, s) || dev_sync(s))) return done(MC_ERR_CUDA);
		std::vector<mc_u64> all;
		if (gather_words(c, mine, 2, all)) return done(MC_ERR_NCCL);
		int64_t carry[2] = {-1, -1};
		for (int r = 0; r < c->comm_rank; r++) for (int k = 0; k < 2; k++) carry[k] = std::max(carry[k], (int64_t)all[(size_t)2 * r + k]);
		if (vbF > 0)
		{
			if (dev_h2d(a.last_nongap + vbF - 1, &carry[0], 8, s) || dev_h2d(a.last_nondup + vbF - 1, &carry[1], 8, s) || dev_sync(s)) return done(MC_ERR_CUDA);
			if (nvo > 0) { device_incmax_i64(a.last_nongap + vbF - 1, nvo + 1, c->d_scan2.p, s); device_incmax_i64(a.last_nondup + vbF - 1, nvo + 1, c->d_scan2.p, s); }
		}
	}
#endif
	for (int64_t t = 0; t < n_tiles; t++) { pack(t); launch_vcscan(a, vb0(t), vb1(t), false, s); }
	if (a.vp.gvcf && nvo > 0) { device_incmax_i64(a.last_normal + vbF, nvo, c->d_scan2.p, s); device_incmax_i64(a.last_event + vbF, nvo, c->d_scan2.p, s); }
	int64_t total = 0;
	if (nvo > 0)
	{
		device_scan_u32(a.cnt + vbF, d_off.as<int64_t>() + vbF, nvo, d_scan.as<int64_t>(), s);
		if (dev_d2h(&total, d_off.as<int64_t>() + vbL, 8, s) || dev_sync(s)) return done(MC_ERR_CUDA);
	}
	mark("count");
	if (d_out.reserve((size_t)(total + 1) * sizeof(mc_variant_rec))) return done(MC_ERR_CUDA);
	a.out = d_out.as<mc_variant_rec>();
	for (int64_t t = 0; t < n_tiles; t++) { pack(t); launch_vcscan(a, vb0(t), vb1(t), true, s); }
	mark("emit");
	// CompByVarPos order on the device (mc_stages_vc.h: vckey_body), then ONE copy of the valid records into page-locked memory
	DBuf &d_vkey = c->d_vc[12], &d_vkey2 = c->d_vc[13], &d_vidx = c->d_vc[14], &d_vidx2 = c->d_vc[15], &d_vsort = c->d_vc[16], &d_out2 = c->d_vc[17];
	mc_u64* d_nvalid = (mc_u64*)d_off.as<int64_t>() + (nvb + 1);   // d_off has nvb + 2 entries; the last one is free
	// sorts `n` record slots of `src` by (gPos, VarType), unused slots last, and leaves the valid ones in d_out2
	auto order = [&](const mc_variant_rec* src, int64_t n, int64_t* n_valid_out) -> int {
		*n_valid_out = 0;
		if (n >= (int64_t)0xFFFFFFFFll) { mc_set_error("mc_variant_scan: more than 2^32 record slots (a monomorphic / gVCF scan of a genome this large does not fit)"); return MC_ERR_OVERFLOW; }
		if (d_vkey.reserve((size_t)(n + 1) * 8) || d_vkey2.reserve((size_t)(n + 1) * 8) || d_vidx.reserve((size_t)(n + 1) * 4) || d_vidx2.reserve((size_t)(n + 1) * 4)
		    || d_vsort.reserve(device_sort_pairs_scratch_bytes(n))) return MC_ERR_CUDA;
		if (dev_zero(d_nvalid, 8, s)) return MC_ERR_CUDA;
		launch_vckey(n, src, d_vkey.as<uint64_t>(), d_vidx.as<uint32_t>(), d_nvalid, s);
		device_sort_pairs(d_vkey.as<uint64_t>(), d_vkey2.as<uint64_t>(), d_vidx.as<uint32_t>(), d_vidx2.as<uint32_t>(), n, d_vsort.p, d_vsort.cap, s);
		int64_t nv = 0;
		if (dev_d2h(&nv, d_nvalid, 8, s) || dev_sync(s)) return MC_ERR_CUDA;
		if (d_out2.reserve((size_t)(nv + 1) * sizeof(mc_variant_rec))) return MC_ERR_CUDA;
		launch_vcgather(nv, src, d_vidx.as<uint32_t>(), d_out2.as<mc_variant_rec>(), s);
		*n_valid_out = nv;
		return MC_OK;
	};
	int64_t n_valid = 0;
	if ((rc = order(a.out, total, &n_valid)) != MC_OK) return done(rc);
#ifndef MC_HOSTEMU
	if (dist)
	{
		// records of all ranks, rank after rank (padded all-gather, compacted, ordered once more: a gap / dup run that ends in this
		// rank's range may have started in an earlier one), and the block depths (equal chunks, in place)
		const int N = c->comm_size;
		mc_u64 mine = (mc_u64)n_valid; std::vector<mc_u64> cnts;
		if (gather_words(c, &mine, 1, cnts)) return done(MC_ERR_NCCL);
		int64_t mx = 1, sum = 0; for (int r = 0; r < N; r++) { mx = std::max(mx, (int64_t)cnts[r]); sum += (int64_t)cnts[r]; }
		DBuf& d_all = d_out;                               // the slots are no longer needed
		if (d_out2.grow_keep((size_t)(mx + 1) * sizeof(mc_variant_rec), (size_t)n_valid * sizeof(mc_variant_rec), s) || d_all.reserve((size_t)(mx + 1) * N * sizeof(mc_variant_rec))) return done(MC_ERR_CUDA);
		if (nccl_fail(ncclAllGather(d_out2.p, d_all.p, (size_t)mx * sizeof(mc_variant_rec), ncclUint8, c->comm, s), "ncclAllGather(variant records)")) return done(MC_ERR_NCCL);
		DBuf& d_cat = c->d_comm_buf;
		if (d_cat.reserve((size_t)(sum + 1) * sizeof(mc_variant_rec))) return done(MC_ERR_CUDA);
		int64_t o = 0;
		for (int r = 0; r < N; r++) { if (cnts[r] && dev_d2d(d_cat.as<mc_variant_rec>() + o, d_all.as<uint8_t>() + (size_t)r * mx * sizeof(mc_variant_rec), (size_t)cnts[r] * sizeof(mc_variant_rec), s)) return done(MC_ERR_CUDA); o += (int64_t)cnts[r]; }
		if ((rc = order(d_cat.as<mc_variant_rec>(), sum, &n_valid)) != MC_OK) return done(rc);
		const int64_t chunk = (c->own_tile / MC_VC_BLOCK);
		if (nccl_fail(ncclAllGather(d_depth.as<int32_t>() + (size_t)c->comm_rank * chunk, d_depth.p, (size_t)chunk, ncclInt32, c->comm, s), "ncclAllGather(block depths)")) return done(MC_ERR_NCCL);
	}
#endif
	if (c->h_vc_out.reserve((size_t)(n_valid + 1) * sizeof(mc_variant_rec)) || c->h_vc_depth.reserve((size_t)(nvb + 1) * 4)) return done(MC_ERR_CUDA);
	if (dev_d2h(c->h_vc_out.p, d_out2.p, (size_t)n_valid * sizeof(mc_variant_rec), s) || dev_d2h(c->h_vc_depth.p, d_depth.p, (size_t)nvb * 4, s) || dev_sync(s)) return done(MC_ERR_CUDA);
	mark("order+d2h");
	mc_variant_rec* vo = c->h_vc_out.as<mc_variant_rec>();
	if (a.vp.gvcf)   // RemoveConsecutiveGenomicVariant (:682-694)
	{
		int64_t w = 0;
		for (int64_t i = 0; i < n_valid; i++)
			if (!(w > 0 && vo[w - 1].VarType == MC_VAR_NOR && vo[i].VarType == MC_VAR_NOR)) vo[w++] = vo[i];
		n_valid = w;
	}
	c->n_vc_out = n_valid;
	*recs = vo; *n_recs = n_valid; *alt_arena = c->ind_seq_out.data();
	*block_depth = c->h_vc_depth.as<int32_t>(); *n_blocks = nvb;
	mark("gvcf");
	return done(MC_OK);
}

// ---- multi-GPU: NCCL over NVLink ------------------------------------------------------------------------
#ifdef MC_HOSTEMU
int mc_comm_unique_id(uint8_t*) { mc_set_error("no NCCL in the developer harness"); return MC_ERR_NCCL; }
int mc_comm_init(mc_ctx*, const uint8_t*, int32_t, int32_t) { mc_set_error("no NCCL in the developer harness"); return MC_ERR_NCCL; }
int mc_profile_allreduce(mc_ctx*, void*) { mc_set_error("no NCCL in the developer harness"); return MC_ERR_NCCL; }
int mc_profile_reduce_scatter(mc_ctx*) { mc_set_error("no NCCL in the developer harness"); return MC_ERR_NCCL; }
#else
int mc_comm_unique_id(uint8_t* out)
{
	if (!out) { mc_set_error("mc_comm_unique_id: null argument"); return MC_ERR_ARG; }
	if (!nccl_api()) return MC_ERR_NCCL;
	ncclUniqueId id;
	if (nccl_fail(ncclGetUniqueId(&id), "ncclGetUniqueId")) return MC_ERR_NCCL;
	memcpy(out, &id, sizeof(id));
	return MC_OK;
}
int mc_comm_init(mc_ctx* c, const uint8_t* id_bytes, int32_t rank, int32_t n_ranks)
{
	if (!c || !id_bytes || rank < 0 || rank >= n_ranks) { mc_set_error("mc_comm_init: bad argument"); return MC_ERR_ARG; }
	if (!nccl_api()) return MC_ERR_NCCL;
	cudaSetDevice(c->prm.device);
	ncclUniqueId id; memcpy(&id, id_bytes, sizeof(id));
	if (nccl_fail(ncclCommInitRank(&c->comm, n_ranks, id, rank), "ncclCommInitRank")) return MC_ERR_NCCL;
	c->comm_rank = rank; c->comm_size = n_ranks;
	return MC_OK;
}

__global__ void mc_seqoff_kernel(mc_indel_rec* r, int64_t n, int32_t add)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) r[i].seq_off += add; }
// independent shards: each 16-bit half clamped to the reference's MaxAlleleCount before the sum (n ranks x 4095 fits 16 bits for n <= 16)
__global__ void mc_clamp_base16_kernel(uint32_t* p, int64_t n)
{
	int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t v = p[i], lo = v & 0xFFFFu, hi = v >> 16;
	if (lo > 4095u || hi > 4095u) p[i] = (lo > 4095u ? 4095u : lo) | ((hi > 4095u ? 4095u : hi) << 16);
}
__global__ void mc_clamp_u8_kernel(uint8_t* p, int64_t n, int hi)
{ int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n && p[i] > hi) p[i] = (uint8_t)hi; }

// Sums what is additive across the shards (difference arrays, base counters, dedup counts, totals) with ncclAllReduce and
// gathers the variable-length records (indels, break points, SV sites), so that afterwards every rank holds the profile
// of the whole library.  `nccl_comm` may be an ncclComm_t created by the caller; NULL uses the one set up by mc_comm_init.
static int profile_reduce(mc_ctx* c, void* nccl_comm, bool scatter)
{
	if (!c) { mc_set_error("mc_profile_allreduce: null context"); return MC_ERR_ARG; }
	if (c->scattered) { mc_set_error("mc_profile_allreduce: the profile has already been reduced (mc_reset starts the next library)"); return MC_ERR_ARG; }
	ncclComm_t comm = nccl_comm ? (ncclComm_t)nccl_comm : c->comm;
	if (!comm) { mc_set_error("mc_profile_allreduce: no communicator (call mc_comm_init first)"); return MC_ERR_NCCL; }
	if (!nccl_api()) return MC_ERR_NCCL;
	if (!c->prm.update_profile) { mc_set_error("mc_profile_allreduce: context was created without update_profile"); return MC_ERR_ARG; }
	cudaSetDevice(c->prm.device);
	cudaStream_t s = c->stream;
	const size_t G = (size_t)c->G;
	int bad = 0;
	const bool dbg = getenv("MC_DEBUG") != nullptr;
	auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t_begin = now();
	ev_record(&c->ev[EV_RED0], s);
	if (c->d_comm_small.reserve(8 * (7 * c->comm_size + 32))) return MC_ERR_CUDA;
	long long* d_tot = c->d_comm_small.as<long long>() + 6 * c->comm_size + 8;
	long long t[5] = {c->tot.total_reads, c->tot.total_mapped, c->tot.total_paired, c->tot.total_distance, c->tot.read_length_sum};
	if (dev_h2d(d_tot, t, sizeof(t), s)) return MC_ERR_CUDA;
	const bool independent = !ordered_mode(c) || comm != c->comm;   // with the ordered exchange readCount and the totals already are the library's
	if (independent)
	{
		// without the global dedup gate every shard may hold up to 15 x 4000 per column: the sum of the 16-bit halves could carry
		// into the neighbouring base and the uint8 readCount sum could wrap; the read-out clamps at 4095 / max_dup anyway
		if (c->comm_size > 16) { mc_set_error("mc_profile_allreduce: independent shards are limited to 16 ranks (packed 16-bit counters)"); return MC_ERR_ARG; }
		mc_clamp_base16_kernel<<<(unsigned)((G * 2 + 255) / 256), 256, 0, s>>>(c->d_base16.as<uint32_t>(), (int64_t)(G * 2));
	}
	// equal tiles of whole MC_TILE_ALIGN units, one per rank (the arrays are padded for it, mc_ctx_create)
	const size_t T = ((G + (size_t)c->comm_size - 1) / (size_t)c->comm_size + MC_TILE_ALIGN - 1) / MC_TILE_ALIGN * MC_TILE_ALIGN, me = (size_t)c->comm_rank;
	if (scatter && (comm != c->comm || c->comm_size > 16 || independent)) { mc_set_error("mc_profile_reduce_scatter: needs the communicator of mc_comm_init with the ordered exchange on (independent shards end with mc_profile_allreduce) and at most 16 ranks"); return MC_ERR_ARG; }
	nccl_api()->GroupStart();
	// packed 2 x uint16 counters are summed as uint32 words: no carry can cross the halves while every column stays below 65536
	if (!scatter)
	{
		bad |= nccl_fail(ncclAllReduce(c->d_base16.p, c->d_base16.p, G * 2, ncclUint32, ncclSum, comm, s), "ncclAllReduce(base16)");
		bad |= nccl_fail(ncclAllReduce(c->d_sdiff.p, c->d_sdiff.p, (G + 1) * 4, ncclInt32, ncclSum, comm, s), "ncclAllReduce(sdiff)");
		bad |= nccl_fail(ncclAllReduce(c->d_cdiff.p, c->d_cdiff.p, G + 1, ncclInt32, ncclSum, comm, s), "ncclAllReduce(cdiff)");
		bad |= nccl_fail(ncclAllReduce(c->d_mdiff.p, c->d_mdiff.p, G + 1, ncclInt32, ncclSum, comm, s), "ncclAllReduce(mdiff)");
		if (independent) bad |= nccl_fail(ncclAllReduce(c->d_rcount.p, c->d_rcount.p, G, ncclUint8, ncclSum, comm, s), "ncclAllReduce(rcount)");
	}
	else
	{
		// in place: rank r receives the sums of tile r where they already lie; half the traffic of the all-reduce, and the
		// read-out that follows (variant scan, summary) runs on N tiles at once
		bad |= nccl_fail(ncclReduceScatter(c->d_base16.p, c->d_base16.as<uint32_t>() + me * T * 2, T * 2, ncclUint32, ncclSum, comm, s), "ncclReduceScatter(base16)");
		bad |= nccl_fail(ncclReduceScatter(c->d_sdiff.p, c->d_sdiff.as<int32_t>() + me * T * 4, T * 4, ncclInt32, ncclSum, comm, s), "ncclReduceScatter(sdiff)");
		bad |= nccl_fail(ncclReduceScatter(c->d_cdiff.p, c->d_cdiff.as<int32_t>() + me * T, T, ncclInt32, ncclSum, comm, s), "ncclReduceScatter(cdiff)");
		bad |= nccl_fail(ncclReduceScatter(c->d_mdiff.p, c->d_mdiff.as<int32_t>() + me * T, T, ncclInt32, ncclSum, comm, s), "ncclReduceScatter(mdiff)");
		if (independent) bad |= nccl_fail(ncclReduceScatter(c->d_rcount.p, c->d_rcount.as<uint8_t>() + me * T, T, ncclUint8, ncclSum, comm, s), "ncclReduceScatter(rcount)");
	}
	if (independent) bad |= nccl_fail(ncclAllReduce(d_tot, d_tot, 5, ncclInt64, ncclSum, comm, s), "ncclAllReduce(totals)");
	bad |= nccl_fail(nccl_api()->GroupEnd(), "ncclGroupEnd");
	if (bad) return MC_ERR_NCCL;
	if (independent) mc_clamp_u8_kernel<<<(unsigned)((G + 255) / 256), 256, 0, s>>>(c->d_rcount.as<uint8_t>(), (int64_t)G, c->prm.max_dup);
	if (scatter)
	{
		c->own_tile = (int64_t)T;
		c->own_beg = (int64_t)std::min(G, me * T); c->own_end = (int64_t)std::min(G, (me + 1) * T);
		for (int k = 0; k < 6; k++) c->own_carry[k] = 0;
		// the difference arrays are read as prefix sums: this rank needs the sums of all columns before its tile = the tile totals of the ranks before it
		DevProfile p; p.base16 = c->d_base16.as<uint32_t>(); p.sdiff = c->d_sdiff.as<int32_t>(); p.cdiff = c->d_cdiff.as<int32_t>(); p.mdiff = c->d_mdiff.as<int32_t>(); p.rcount = c->d_rcount.as<uint8_t>();
		const int64_t nb = ((int64_t)G + MC_PROF_BLOCK - 1) / MC_PROF_BLOCK;
		DBuf& d_sums = c->d_vc[0];
		if (d_sums.reserve((size_t)(6 * nb + 8) * 8) || c->d_scan2.reserve(device_scan_scratch_bytes(nb)) || dev_zero(d_sums.as<int64_t>() + 6 * nb, 48, s)) return MC_ERR_CUDA;
		mc_u64 mine[6] = {0, 0, 0, 0, 0, 0};
		if (c->own_end > c->own_beg && (profile_prefix(c, p, nb, d_sums.as<int64_t>(), s) || dev_d2h(mine, d_sums.as<int64_t>() + 6 * nb, 48, s) || dev_sync(s))) return MC_ERR_CUDA;
		std::vector<mc_u64> all;
		if (gather_words(c, mine, 6, all)) return MC_ERR_NCCL;
		for (int r = 0; r < c->comm_rank; r++) for (int k = 0; k < 6; k++) c->own_carry[k] += (int64_t)all[(size_t)6 * r + k];
		c->scattered = true;
	}
	if (dev_d2h(t, d_tot, sizeof(t), s)) return MC_ERR_CUDA;   // completes with the next synchronize (record exchange below)
	// variable-length records (break points, indels with their sequences, SV sites): every rank ends up with the records of
	// all ranks in rank (= file) order.  The payload goes device to device; only the counts pass through the host.
	const int n = c->comm_size;
	long long* d_cnt = c->d_comm_small.as<long long>();                 // [n][4] persistent cursors, then [n][2] site counts, then mine
	long long my_sites[2] = {(long long)c->inv_sites.size(), (long long)c->tnl_sites.size()};
	if (c->d_comm_small.cap < (size_t)8 * (6 * n + 16)) { mc_set_error("mc_profile_allreduce: internal error (scratch)"); return MC_ERR_CUDA; }
	if (dev_h2d(d_cnt + 6 * n, my_sites, 16, s)) return MC_ERR_CUDA;
	nccl_api()->GroupStart();
	bad |= nccl_fail(ncclAllGather(c->d_pbump.p, d_cnt, 4, ncclInt64, comm, s), "ncclAllGather(record counts)");
	bad |= nccl_fail(ncclAllGather(d_cnt + 6 * n, d_cnt + 4 * n, 2, ncclInt64, comm, s), "ncclAllGather(site counts)");
	bad |= nccl_fail(nccl_api()->GroupEnd(), "ncclGroupEnd");
	if (bad) return MC_ERR_NCCL;
	std::vector<long long> cnt((size_t)6 * n);
	if (dev_d2h(cnt.data(), d_cnt, (size_t)8 * 6 * n, s) || dev_sync(s)) return MC_ERR_CUDA;
	const double t_reduced = now();
	if (independent)
	{
		c->tot.total_reads = t[0]; c->tot.total_mapped = t[1]; c->tot.total_paired = t[2]; c->tot.total_distance = t[3]; c->tot.read_length_sum = t[4];
		if (c->tot.total_paired > 1000) c->tot.avg_dist = (uint32_t)(int)(1. * c->tot.total_distance / c->tot.total_paired + .5);
	}
	size_t mx[4] = {0, 0, 0, 0}, tot[4] = {0, 0, 0, 0};                  // 0 break points, 1 indel records, 2 indel sequences, 3 sites (inv then tnl)
	const size_t esz[4] = {8, sizeof(mc_indel_rec), 1, sizeof(mc_site_rec)};
	auto count_of = [&](int r, int k) -> size_t { return k < 3 ? (size_t)cnt[4 * r + k] : (size_t)(cnt[4 * n + 2 * r] + cnt[4 * n + 2 * r + 1]); };
	for (int r = 0; r < n; r++) for (int k = 0; k < 4; k++) { mx[k] = std::max(mx[k], count_of(r, k)); tot[k] += count_of(r, k); }
	for (int k = 0; k < 4; k++) mx[k] = ((mx[k] * esz[k] + 15) & ~(size_t)15);     // padded bytes per rank
	if (tot[2] >= 0x7fffffffull) { mc_set_error("mc_profile_allreduce: merged indel sequences exceed 2 GiB"); return MC_ERR_OVERFLOW; }
	// sources must be readable for the padded length
	const size_t my_bp = (size_t)cnt[4 * c->comm_rank], my_ind = (size_t)cnt[4 * c->comm_rank + 1], my_seq = (size_t)cnt[4 * c->comm_rank + 2];
	bad = c->d_bp.grow_keep(mx[0] + 64, my_bp * 8, s) || c->d_ind.grow_keep(mx[1] + 64, my_ind * sizeof(mc_indel_rec), s) || c->d_ind_seq.grow_keep(mx[2] + 64, my_seq, s);
	std::vector<mc_site_rec> sites(c->inv_sites); sites.insert(sites.end(), c->tnl_sites.begin(), c->tnl_sites.end());
	size_t goff[5] = {0, 0, 0, 0, 0};
	for (int k = 0; k < 4; k++) goff[k + 1] = goff[k] + mx[k] * n;
	bad = bad || c->d_comm_buf.reserve(goff[4] + mx[3] + 64) || c->h_comm_buf.reserve(mx[3] * n + 64);
	if (bad) return MC_ERR_CUDA;
	uint8_t* gb = c->d_comm_buf.as<uint8_t>(); uint8_t* d_my_sites = gb + goff[4];
	if (dev_h2d(d_my_sites, sites.data(), sites.size() * sizeof(mc_site_rec), s)) return MC_ERR_CUDA;
	nccl_api()->GroupStart();
	if (mx[0]) bad |= nccl_fail(ncclAllGather(c->d_bp.p, gb + goff[0], mx[0], ncclUint8, comm, s), "ncclAllGather(break points)");
	if (mx[1]) bad |= nccl_fail(ncclAllGather(c->d_ind.p, gb + goff[1], mx[1], ncclUint8, comm, s), "ncclAllGather(indel records)");
	if (mx[2]) bad |= nccl_fail(ncclAllGather(c->d_ind_seq.p, gb + goff[2], mx[2], ncclUint8, comm, s), "ncclAllGather(indel sequences)");
	if (mx[3]) bad |= nccl_fail(ncclAllGather(d_my_sites, gb + goff[3], mx[3], ncclUint8, comm, s), "ncclAllGather(sites)");
	bad |= nccl_fail(nccl_api()->GroupEnd(), "ncclGroupEnd");
	if (bad) return MC_ERR_NCCL;
	// compact the padded segments back into the persistent arrays, rank after rank
	bad = c->d_bp.reserve(tot[0] * 8 + 64) || c->d_ind.reserve(tot[1] * sizeof(mc_indel_rec) + 64) || c->d_ind_seq.reserve(tot[2] + 64);
	if (bad) return MC_ERR_CUDA;
	size_t o_bp = 0, o_ind = 0, o_seq = 0;
	for (int r = 0; r < n; r++)
	{
		const size_t nb = count_of(r, 0), ni = count_of(r, 1), ns = count_of(r, 2);
		bad |= nb ? dev_d2d(c->d_bp.as<uint8_t>() + o_bp * 8, gb + goff[0] + mx[0] * r, nb * 8, s) : 0;
		bad |= ni ? dev_d2d(c->d_ind.as<uint8_t>() + o_ind * sizeof(mc_indel_rec), gb + goff[1] + mx[1] * r, ni * sizeof(mc_indel_rec), s) : 0;
		bad |= ns ? dev_d2d(c->d_ind_seq.as<uint8_t>() + o_seq, gb + goff[2] + mx[2] * r, ns, s) : 0;
		if (ni && o_seq) mc_seqoff_kernel<<<(unsigned)((ni + 255) / 256), 256, 0, s>>>(c->d_ind.as<mc_indel_rec>() + o_ind, (int64_t)ni, (int32_t)o_seq);
		o_bp += nb; o_ind += ni; o_seq += ns;
	}
	PersistBumps pb; pb.bp = tot[0]; pb.ind = tot[1]; pb.ind_seq = tot[2]; pb.pad = 0;
	bad = bad || dev_h2d(c->d_pbump.p, &pb, sizeof(pb), s) || dev_d2h(c->h_comm_buf.p, gb + goff[3], mx[3] * n, s) || dev_sync(s);
	if (bad) return MC_ERR_CUDA;
	c->bp_cap = c->d_bp.cap / 8; c->ind_cap = c->d_ind.cap / sizeof(mc_indel_rec); c->ind_seq_cap = c->d_ind_seq.cap;
	std::vector<mc_site_rec> inv, tnl;
	for (int r = 0; r < n; r++)
	{
		const mc_site_rec* p = (const mc_site_rec*)(c->h_comm_buf.as<uint8_t>() + mx[3] * r);
		inv.insert(inv.end(), p, p + cnt[4 * n + 2 * r]); tnl.insert(tnl.end(), p + cnt[4 * n + 2 * r], p + cnt[4 * n + 2 * r] + cnt[4 * n + 2 * r + 1]);
	}
	c->inv_sites = inv; c->tnl_sites = tnl;
	ev_record(&c->ev[EV_RED1], s);
	if (dev_sync(s)) return MC_ERR_CUDA;
	c->stats.ms_reduce += ev_ms(&c->ev[EV_RED0], &c->ev[EV_RED1]);
	const size_t mine_bytes = my_bp * 8 + my_ind * sizeof(mc_indel_rec) + my_seq;
	if (dbg) fprintf(stderr, "[mc] rank %d %s: counters %.3f ms, records %.3f ms (%zu bytes mine)\n", c->comm_rank, scatter ? "reduce-scatter" : "allreduce", t_reduced - t_begin, now() - t_reduced, mine_bytes);
	return MC_OK;
}
int mc_profile_allreduce(mc_ctx* c, void* nccl_comm) { return profile_reduce(c, nccl_comm, false); }
// The scalable end of a library on N GPUs: every rank keeps the sums of ONE tile of the genome (ncclReduceScatter in place, half
// the traffic of the all-reduce) and the read-out calls become collectives over the tiles - mc_variant_scan scans the tile
// and gathers the records of all ranks, mc_profile_summary / mc_profile_checksum combine the tiles' partial results,
// mc_profile_read serves the tile's columns (mc_profile_owned tells which).  Indel / break-point / SV-site records are
// gathered to every rank as in mc_profile_allreduce.
int mc_profile_reduce_scatter(mc_ctx* c) { return profile_reduce(c, nullptr, true); }
#endif
int mc_profile_owned(const mc_ctx* c, int64_t* beg, int64_t* end)
{
	if (!c || !beg || !end) { mc_set_error("mc_profile_owned: null argument"); return MC_ERR_ARG; }
	*beg = c->own_beg; *end = c->own_end;
	return MC_OK;
}

// Operator-level entry: seeding, locate and clustering of a batch (the front of run_batch), results back to the host.
int mc_seed_cluster_batch(mc_ctx* c, const mc_batch_in* in, mc_seed_cluster_out* out)
{
	if (!c || !in || !out) { mc_set_error("mc_seed_cluster_batch: null argument"); return MC_ERR_ARG; }
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	memset(out, 0, sizeof(*out));
	const mc_stream_t s = c->stream;
	const int32_t paired_saved = c->prm.paired;
	c->prm.paired = 0;                                   // staging must not insist on an even number of reads
	int rc = stage_reads(c, in, c->cur, s, 1);
	c->prm.paired = paired_saved;
	if (rc) return rc;
	Staged& st = c->cur;
	const int64_t n = st.n_reads;
	out->n_reads = n;
	if (n == 0) return MC_OK;
	PipeArgs a; memset(&a, 0, sizeof(a));
	a.ix = c->ix; a.pr.paired = 0; a.pr.max_pos_diff = c->prm.max_pos_diff;
	a.st = c->d_stats.as<DevStats>(); a.n_reads = n; a.seq = st.seq.as<uint8_t>() - st.base; a.roff = st.roff.as<int64_t>(); a.seed_off = st.seed_off.as<int64_t>();
	a.n_slots = st.n_slots;
	int bad = c->d_slot_freq.reserve(st.n_slots * 4) || c->d_seeds.reserve(st.n_slots * sizeof(Seed)) || c->d_slot_loc.reserve((st.n_slots + 1) * 8) || c->d_scan.reserve(device_scan_scratch_bytes(st.n_slots));
	bad |= c->d_cand_off.reserve((n + 1) * 4) || c->d_rflag.reserve(n + 1) || c->d_npair.reserve(n * 4) || c->d_ncand0.reserve(n * 4);
	if (bad) return MC_ERR_CUDA;
	a.slot_freq = c->d_slot_freq.as<uint32_t>(); a.seeds = c->d_seeds.as<Seed>(); a.slot_loc = c->d_slot_loc.as<int64_t>();
	a.cand_off = c->d_cand_off.as<int32_t>(); a.rflag = c->d_rflag.as<uint8_t>(); a.npair = c->d_npair.as<int32_t>(); a.ncand0 = c->d_ncand0.as<int32_t>();
	bad |= dev_zero(c->d_slot_freq.p, st.n_slots * 4, s) || dev_zero(c->d_stats.p, sizeof(DevStats) - 2 * sizeof(mc_u64), s) || dev_d2d(&c->d_stats.as<DevStats>()->overflow, st.flag.p, 8, s);
	launch_seed(a, 0, n, s);
	device_scan_u32(a.slot_freq, c->d_slot_loc.as<int64_t>(), st.n_slots, c->d_scan.as<int64_t>(), s);
	int64_t* h_small = c->h_small.as<int64_t>();
	bad |= dev_d2h(h_small, c->d_slot_loc.as<int64_t>() + st.n_slots, 8, s) || dev_sync(s);
	if (bad) return MC_ERR_CUDA;
	const int64_t n_locs = h_small[0];
	a.n_locs = n_locs;
	if (n_locs >= 0x7fffffffll) { mc_set_error("mc_seed_cluster_batch: batch too large; split it"); return MC_ERR_ARG; }
	bad |= c->d_pairs.reserve((size_t)(n_locs + 1) * sizeof(SPair)) || c->d_cands.reserve((size_t)(n_locs + 1) * sizeof(Cand));
	if (bad) return MC_ERR_CUDA;
	a.pairs = c->d_pairs.as<SPair>(); a.pair_cap = n_locs; a.cands = c->d_cands.as<Cand>();
	launch_expand(a, st.n_slots, s);
	launch_locate(a, n_locs, s);
	launch_cluster(a, n, s);
	// back to the host: the per-read offsets are slot_loc[seed_off[r]] (pairs and, in single-end layout, clusters alike)
	std::vector<int64_t> seed_off((size_t)n + 1), slot_loc((size_t)st.n_slots + 1);
	c->sc_pair_off.assign((size_t)n + 1, 0); c->sc_npairs.resize((size_t)n); c->sc_nclusters.resize((size_t)n);
	c->sc_pairs.resize((size_t)n_locs + 1); c->sc_clusters.resize((size_t)n_locs + 1);
	mc_u64 flag = 0;
	bad |= dev_d2h(seed_off.data(), st.seed_off.p, (size_t)(n + 1) * 8, s) || dev_d2h(slot_loc.data(), c->d_slot_loc.p, (size_t)(st.n_slots + 1) * 8, s);
	bad |= dev_d2h(c->sc_npairs.data(), c->d_npair.p, (size_t)n * 4, s) || dev_d2h(c->sc_nclusters.data(), c->d_ncand0.p, (size_t)n * 4, s);
	bad |= dev_d2h(c->sc_pairs.data(), c->d_pairs.p, (size_t)n_locs * sizeof(SPair), s) || dev_d2h(c->sc_clusters.data(), c->d_cands.p, (size_t)n_locs * sizeof(Cand), s);
	bad |= dev_d2h(&flag, &c->d_stats.as<DevStats>()->overflow, 8, s) || dev_sync(s);
	if (bad) return MC_ERR_CUDA;
	if (flag) { mc_set_error("mc_seed_cluster_batch: a read is longer than %d bases (or the offsets decrease)", MC_MAX_RLEN); return MC_ERR_ARG; }
	for (int64_t r = 0; r <= n; r++) c->sc_pair_off[(size_t)r] = slot_loc[(size_t)seed_off[(size_t)r]];
	out->pair_off = c->sc_pair_off.data(); out->n_pairs = c->sc_npairs.data(); out->pairs = (const mc_simple_pair*)c->sc_pairs.data();
	out->cluster_off = c->sc_pair_off.data(); out->n_clusters = c->sc_nclusters.data(); out->clusters = (const mc_cluster*)c->sc_clusters.data();
	return MC_OK;
}

// Operator-level entry: n independent BWT_Search queries through the extension / locate primitives the pipeline uses.
int mc_bwt_search_batch(mc_ctx* c, int64_t n, const uint8_t* codes, const int64_t* off, const int32_t* start, int32_t* out_len, int32_t* out_freq, uint64_t* out_loc)
{
	if (!c || n < 0 || (n > 0 && (!codes || !off || !start || !out_len || !out_freq || !out_loc))) { mc_set_error("mc_bwt_search_batch: bad argument"); return MC_ERR_ARG; }
	if (n == 0) return MC_OK;
#ifndef MC_HOSTEMU
	cudaSetDevice(c->prm.device);
#endif
	const mc_stream_t s = c->stream;
	const int64_t bytes = off[n] - off[0];
	if (bytes < 0) { mc_set_error("mc_bwt_search_batch: offsets decrease"); return MC_ERR_ARG; }
	DBuf d_codes, d_off, d_start, d_len, d_freq, d_loc;
	int bad = d_codes.reserve(bytes + 16) || d_off.reserve((n + 1) * 8) || d_start.reserve(n * 4) || d_len.reserve(n * 4) || d_freq.reserve(n * 4) || d_loc.reserve(n * MC_MAX_OCC * 8);
	bad = bad || dev_h2d(d_codes.p, codes + off[0], bytes, s) || dev_h2d(d_off.p, off, (n + 1) * 8, s) || dev_h2d(d_start.p, start, n * 4, s) || dev_zero(d_loc.p, n * MC_MAX_OCC * 8, s);
	if (!bad)
	{
		SearchArgs a; a.ix = c->ix; a.codes = d_codes.as<uint8_t>() - off[0]; a.off = d_off.as<int64_t>(); a.start = d_start.as<int32_t>();
		a.len = d_len.as<int32_t>(); a.freq = d_freq.as<int32_t>(); a.loc = d_loc.as<uint64_t>();
		launch_bwtsearch(a, n, s);
		bad = dev_d2h(out_len, d_len.p, n * 4, s) || dev_d2h(out_freq, d_freq.p, n * 4, s) || dev_d2h(out_loc, d_loc.p, n * MC_MAX_OCC * 8, s) || dev_sync(s);
	}
	DBuf* all[] = {&d_codes, &d_off, &d_start, &d_len, &d_freq, &d_loc};
	for (DBuf* b : all) b->release();
	if (bad) { mc_set_error("mc_bwt_search_batch: CUDA failure"); return MC_ERR_CUDA; }
	return MC_OK;
}
// Operator-level entry: n independent gapped fills through dp_body (the kernel the pipeline uses).
int mc_align_batch(mc_ctx* c, int32_t use_ksw2, int64_t n, const uint8_t* s1, const int64_t* off1, const uint8_t* s2, const int64_t* off2,
                   const int64_t* out_off, uint8_t* out1, uint8_t* out2, int32_t* out_len)
{
#ifndef MC_HOSTEMU
	if (c) cudaSetDevice(c->prm.device);   // a process may hold contexts on several GPUs
#endif
	if (!c || n < 0 || (n > 0 && (!s1 || !off1 || !s2 || !off2 || !out_off || !out1 || !out2 || !out_len))) { mc_set_error("mc_align_batch: bad argument"); return MC_ERR_ARG; }
	if (n == 0) return MC_OK;
	const mc_stream_t s = c->stream;
	std::vector<mc_frag_out> fr((size_t)n); std::vector<DpTask> tk((size_t)n);
	int64_t aln_bytes = 0, ws_bytes = 0;
	for (int64_t i = 0; i < n; i++)
	{
		const int64_t m = off1[i + 1] - off1[i], g = off2[i + 1] - off2[i];
		if (m <= 0 || g <= 0 || m > MC_MAX_RLEN || g > 2 * MC_MAX_RLEN || out_off[i + 1] - out_off[i] < m + g) { mc_set_error("mc_align_batch: problem %lld has bad sizes", (long long)i); return MC_ERR_ARG; }
		mc_frag_out& f = fr[i]; memset(&f, 0, sizeof(f));
		f.rLen = (int32_t)m; f.gLen = (int32_t)g; f.aln_cap = (int32_t)(m + g); f.aln_off = (int32_t)aln_bytes; aln_bytes += 2 * (m + g);
		DpTask& t = tk[i]; t.frag = (int32_t)i; t.m = (int32_t)m; t.n = (int32_t)g; t.pad = 0; t.ws_off = ws_bytes; ws_bytes += dp_ws_bytes((int)m, (int)g);
		if (aln_bytes >= 0x7fffffffll) { mc_set_error("mc_align_batch: too many bases in one call"); return MC_ERR_ARG; }
	}
	std::vector<uint8_t> h_aln((size_t)aln_bytes);
	for (int64_t i = 0; i < n; i++)
	{
		memcpy(h_aln.data() + fr[i].aln_off, s1 + off1[i], (size_t)fr[i].rLen);
		memcpy(h_aln.data() + fr[i].aln_off + fr[i].aln_cap, s2 + off2[i], (size_t)fr[i].gLen);
	}
	DBuf d_fr, d_tk, d_aln, d_ws, d_cnt;
	mc_u64 h_cnt = (mc_u64)n;
	int bad = d_fr.reserve(n * sizeof(mc_frag_out)) || d_tk.reserve(n * sizeof(DpTask)) || d_aln.reserve(aln_bytes) || d_ws.reserve(ws_bytes) || d_cnt.reserve(8);
	bad = bad || dev_h2d(d_cnt.p, &h_cnt, 8, s);
	bad = bad || dev_h2d(d_fr.p, fr.data(), n * sizeof(mc_frag_out), s) || dev_h2d(d_tk.p, tk.data(), n * sizeof(DpTask), s) || dev_h2d(d_aln.p, h_aln.data(), aln_bytes, s);
	if (!bad)
	{
		PipeArgs a; memset(&a, 0, sizeof(a));
		a.pr.alg_ksw2 = use_ksw2; a.st = c->d_stats.as<DevStats>(); a.frags = d_fr.as<mc_frag_out>(); a.tasks = d_tk.as<DpTask>(); a.aln = d_aln.as<uint8_t>(); a.dpws = d_ws.as<uint8_t>();
		a.task_bump = d_cnt.as<mc_u64>(); a.task_cap = n; a.task_begin = 0;
		launch_dp(a, n, s);
		bad = dev_d2h(fr.data(), d_fr.p, n * sizeof(mc_frag_out), s) || dev_d2h(h_aln.data(), d_aln.p, aln_bytes, s) || dev_sync(s);
	}
	d_fr.release(); d_tk.release(); d_aln.release(); d_ws.release(); d_cnt.release();
	if (bad) return MC_ERR_CUDA;
	for (int64_t i = 0; i < n; i++)
	{
		out_len[i] = fr[i].aln_len;
		memcpy(out1 + out_off[i], h_aln.data() + fr[i].aln_off, (size_t)fr[i].aln_len);
		memcpy(out2 + out_off[i], h_aln.data() + fr[i].aln_off + fr[i].aln_cap, (size_t)fr[i].aln_len);
	}
	return MC_OK;
}

} // extern "C"
