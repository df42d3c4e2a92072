// FM-index construction, load and save (host side of libmapcaller_b200.so).
//
// Replaces bwa_idx_build / bwa_idx_load / RestoreReferenceInfo of the reference
// (src/BWT_Index/bwtindex.c:77, src/bwt_index.cpp:150,232).  The reference grows the BWT
// incrementally (BWT-SW, src/BWT_Index/bwt_gen.c); the BWT of a text is unique, so this builder is
// free to get there differently: it sorts the suffixes of  fwd + revcomp  directly (a 4^K bucket pass
// followed by multi-threaded per-bucket comparison sorts over the 2-bit packed text) and then emits
// the reference's on-disk / in-memory layout:
//   .bwt  primary, L2[1..4], then per 128 symbols: 4 x uint64 running counts + 8 x uint32 packed
//         symbols (16 per word, first symbol in the top bits), plus one trailing count record
//         (src/BWT_Index/bwtindex.c:53-75, bwt.c:174-183)
//   .sa   primary, L2[1..4], sa_intv, seq_len, then the SA value of every 32nd ROW, row 0 excluded
//         (bwt.c:101-123,185-196); row 0 is the empty suffix and reads back as (uint64)-1
//   .pac  forward-only 2-bit text + length trailer (bntseq.c:170-211); .ann/.amb text (bntseq.c:59-91)
#include "../../include/mapcaller_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <cstdarg>
#include <stdexcept>
#include <string>
#include <thread>
#include <functional>
#include <array>
#include <vector>
#include <chrono>

void mc_set_error(const char* fmt, ...);
#ifndef MC_HOSTEMU
// index_gpu.cu: BWT symbols by row, `primary` and the sampled suffix array of the packed text, sorted in chunks on the GPU
extern "C" int mc_gpu_bwt_build(const uint64_t* words, size_t n_words, int64_t n, int device, uint64_t* rowsym_out, uint64_t* primary_out, uint64_t* samples_out);
#endif

struct mc_index {
	std::vector<uint32_t> bwt_store;
	std::vector<uint64_t> sa_store;
	std::vector<uint8_t> pac_store;
	std::vector<int32_t> chrom_len;
	std::vector<std::string> chrom_name, chrom_anno;
	std::vector<const char*> chrom_name_ptr;
	struct Hole { int64_t offset; int32_t len; char amb; };
	std::vector<Hole> holes;
	std::vector<int32_t> chrom_n_ambs;
	mc_index_view v;
};

namespace {

const unsigned char kNt4[256] = {
#define R16 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4
	R16, R16, R16, R16,
	4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
	4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
	R16, R16, R16, R16, R16, R16, R16, R16
#undef R16
};

// 2-bit packed text, 32 bases per word, first base in the top bits; two zero words of padding
struct PackedText {
	std::vector<uint64_t> w;
	int64_t n;
	inline uint64_t window(int64_t i) const
	{
		int64_t k = i >> 5; int s = (int)(i & 31) << 1;
		return s ? (w[k] << s) | (w[k + 1] >> (64 - s)) : w[k];
	}
	inline int base(int64_t i) const { return (int)(w[i >> 5] >> ((~i & 31) << 1)) & 3; }
};

// suffix a < suffix b ; `skip` leading bases are known equal
struct SuffixLess {
	const PackedText* t; int skip;
	bool operator()(uint64_t a, uint64_t b) const
	{
		const int64_t n = t->n;
		int64_t pa = (int64_t)a + skip, pb = (int64_t)b + skip;
		for (;;)
		{
			int64_t ra = n - pa, rb = n - pb;
			if (ra <= 0 || rb <= 0) return ra < rb;
			uint64_t wa = t->window(pa), wb = t->window(pb);
			int64_t m = std::min<int64_t>(32, std::min(ra, rb));
			if (m < 32) { uint64_t mask = ~0ull << ((32 - m) << 1); wa &= mask; wb &= mask; }
			if (wa != wb) return wa < wb;
			if (m < 32) return ra < rb; // the suffix that ran out is the smaller one
			pa += 32; pb += 32;
		}
	}
};

template <class IdxT>
void sort_suffixes(const PackedText& t, std::vector<IdxT>& sa, int n_threads)
{
	const int64_t n = t.n;
	int K = 1; while (K < 12 && (1ll << (2 * (K + 1))) <= n / 4) K++;
	const int64_t nb = 1ll << (2 * K);
	std::vector<int64_t> cnt(nb + 1, 0);
	auto key_at = [&](int64_t i) -> int64_t { return (int64_t)(t.window(i) >> (64 - 2 * K)); }; // zero padded past the end
	for (int64_t i = 0; i < n; i++) cnt[key_at(i) + 1]++;
	for (int64_t b = 0; b < nb; b++) cnt[b + 1] += cnt[b];
	sa.resize(n);
	{
		std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
		for (int64_t i = 0; i < n; i++) sa[cur[key_at(i)]++] = (IdxT)i;
	}
	std::atomic<int64_t> next(0);
	const int64_t grain = std::max<int64_t>(1, nb / (n_threads * 64));
	auto worker = [&]() {
		SuffixLess less{&t, 0};
		for (;;)
		{
			int64_t b0 = next.fetch_add(grain); if (b0 >= nb) break;
			int64_t b1 = std::min(nb, b0 + grain);
			for (int64_t b = b0; b < b1; b++)
			{
				int64_t lo = cnt[b], hi = cnt[b + 1];
				if (hi - lo > 1) std::sort(sa.begin() + lo, sa.begin() + hi, [&](IdxT x, IdxT y) { return less((uint64_t)x, (uint64_t)y); });
			}
		}
	};
	std::vector<std::thread> th;
	for (int i = 1; i < n_threads; i++) th.emplace_back(worker);
	worker();
	for (auto& x : th) x.join();
}

void finish_view(mc_index* ix, uint64_t primary, const uint64_t L2[5], uint64_t seq_len, int64_t G)
{
	ix->chrom_name_ptr.clear();
	for (auto& s : ix->chrom_name) ix->chrom_name_ptr.push_back(s.c_str());
	mc_index_view& v = ix->v;
	v.bwt = ix->bwt_store.data(); v.bwt_size = ix->bwt_store.size();
	v.primary = primary; for (int i = 0; i < 5; i++) v.L2[i] = L2[i];
	v.seq_len = seq_len; v.sa = ix->sa_store.data(); v.n_sa = ix->sa_store.size(); v.sa_intv = 32;
	v.pac = ix->pac_store.data(); v.genome_size = G;
	v.n_chrom = (int32_t)ix->chrom_len.size(); v.chrom_len = ix->chrom_len.data(); v.chrom_name = ix->chrom_name_ptr.data();
}

static thread_local int g_sort_device = -1;   // >= 0: mc_index_build_gpu is running, sort the suffixes on that device

template <class IdxT>
int build_core(mc_index* ix, const uint8_t* fwd, int64_t G, int n_threads)
{
	const int64_t N = 2 * G;
	const bool verbose = getenv("MC_DEBUG") != nullptr;
	auto t_last = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		const auto now = std::chrono::steady_clock::now();
		if (verbose) fprintf(stderr, "[mc] index build: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
		t_last = now;
	};
	PackedText t; t.n = N; t.w.assign((N + 31) / 32 + 2, 0);
	if (n_threads < 1) n_threads = 1;
	// every helper below cuts its index range into n_threads contiguous pieces
	auto parallel = [&](int64_t n_items, const std::function<void(int, int64_t, int64_t)>& fn) {
		const int T = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n_items / 4096));
		std::vector<std::thread> th;
		for (int k = 0; k < T; k++)
		{
			const int64_t b = n_items * k / T, e = n_items * (k + 1) / T;
			if (k + 1 < T) th.emplace_back(fn, k, b, e); else fn(k, b, e);
		}
		for (auto& x : th) x.join();
	};
	// packed text: word k holds positions 32k .. 32k+31 of forward + reverse complement (one writer per word)
	uint64_t L2[5] = {0, 0, 0, 0, 0};
	{
		std::vector<std::array<uint64_t, 4> > cnt((size_t)n_threads, std::array<uint64_t, 4>{{0, 0, 0, 0}});
		parallel((N + 31) / 32, [&](int tid, int64_t w0, int64_t w1) {
			std::array<uint64_t, 4> c{{0, 0, 0, 0}};
			for (int64_t k = w0; k < w1; k++)
			{
				uint64_t word = 0;
				const int64_t p1 = std::min<int64_t>(N, 32 * k + 32);
				for (int64_t p = 32 * k; p < p1; p++)
				{
					const int sym = p < G ? (fwd[p] & 3) : 3 - (fwd[N - 1 - p] & 3);
					word |= (uint64_t)sym << ((~p & 31) << 1); c[sym]++;
				}
				t.w[k] = word;
			}
			cnt[tid] = c;
		});
		for (auto& c : cnt) for (int k = 0; k < 4; k++) L2[k + 1] += c[k];
	}
	for (int c = 0; c < 4; c++) L2[c + 1] += L2[c];
	lap("packed text");

	// rows: 0 = empty suffix, r >= 1 = sa[r-1].  BWT symbol of row r = text[pos-1]; the row with pos == 0 is `primary`.
	std::vector<IdxT> sa;
	PackedText rowsym; rowsym.n = 0;                 // GPU path: symbol of row S + 1 at S, no suffix array on the host at all
	uint64_t primary = 0;
	const uint64_t n_sa = ((uint64_t)N + 32) / 32;
	ix->sa_store.assign(n_sa, 0);
#ifndef MC_HOSTEMU
	if (g_sort_device >= 0)
	{
		rowsym.n = N; rowsym.w.assign((N + 31) / 32, 0);
		int rc = mc_gpu_bwt_build(t.w.data(), t.w.size(), N, g_sort_device, rowsym.w.data(), &primary, ix->sa_store.data());
		if (rc != MC_OK) return rc;
	}
	else
#endif
	{
		sort_suffixes<IdxT>(t, sa, n_threads);
		std::vector<uint64_t> found((size_t)n_threads, 0);
		parallel(N, [&](int tid, int64_t r0, int64_t r1) { for (int64_t r = r0; r < r1; r++) if (sa[r] == 0) found[tid] = (uint64_t)r + 1; });
		for (uint64_t f : found) if (f) primary = f;
		ix->sa_store[0] = (uint64_t)-1;
		for (uint64_t j = 1; j < n_sa; j++) ix->sa_store[j] = (uint64_t)sa[j * 32 - 1];
	}

	lap(g_sort_device >= 0 ? "rows sorted (GPU)" : "suffixes sorted (host)");
	const uint64_t n_occ = (uint64_t)(N + 127) / 128 + 1;
	ix->bwt_store.assign(((uint64_t)(N + 15) >> 4) + n_occ * 8, 0);
	uint32_t* out = ix->bwt_store.data();
	// the row that holds the whole text ($ in its BWT column) is skipped: symbol k of the BWT string belongs to row k
	// below it and to row k + 1 from it on
	const bool from_rows = rowsym.n > 0;
	auto symbol = [&](uint64_t k) -> int {
		if (k == 0) return t.base(N - 1);
		const uint64_t r = k < primary ? k : k + 1;
		return from_rows ? rowsym.base((int64_t)r - 1) : t.base((int64_t)sa[r - 1] - 1);
	};
	// blocks of 128 symbols: 8 words of running counts, then 8 words of symbols; a last record of counts closes the string.
	// Two passes over contiguous block ranges: symbol counts per range, then the ranges are written with their start counts.
	const int64_t n_blocks = (N + 127) / 128;
	{
		std::vector<std::array<uint64_t, 4> > cnt((size_t)n_threads, std::array<uint64_t, 4>{{0, 0, 0, 0}});
		std::vector<std::pair<int64_t, int64_t> > range((size_t)n_threads, std::make_pair((int64_t)0, (int64_t)0));
		parallel(n_blocks, [&](int tid, int64_t b0, int64_t b1) {
			std::array<uint64_t, 4> c{{0, 0, 0, 0}};
			const uint64_t k1 = std::min<uint64_t>((uint64_t)N, (uint64_t)b1 * 128);
			for (uint64_t k = (uint64_t)b0 * 128; k < k1; k++) c[symbol(k)]++;
			cnt[tid] = c; range[tid] = std::make_pair(b0, b1);
		});
		std::vector<std::array<uint64_t, 4> > start((size_t)n_threads);
		uint64_t run[4] = {0, 0, 0, 0};
		for (int k = 0; k < n_threads; k++) { for (int c = 0; c < 4; c++) { start[k][c] = run[c]; run[c] += cnt[k][c]; } }
		parallel(n_blocks, [&](int tid, int64_t b0, int64_t b1) {
			uint64_t c4[4] = {start[tid][0], start[tid][1], start[tid][2], start[tid][3]};
			if (range[tid].first != b0 || range[tid].second != b1) return;   // cannot happen: same cut as the counting pass
			for (int64_t b = b0; b < b1; b++)
			{
				uint32_t* rec = out + 16 * b;
				memcpy(rec, c4, 32);
				const uint64_t k1 = std::min<uint64_t>((uint64_t)N, (uint64_t)b * 128 + 128);
				for (uint64_t k = (uint64_t)b * 128; k < k1; k++)
				{
					const int sym = symbol(k);
					rec[8 + ((k & 127) >> 4)] |= (uint32_t)sym << ((~k & 15) << 1);
					c4[sym]++;
				}
			}
		});
		const uint64_t w = 8 * (uint64_t)n_blocks + ((uint64_t)(N + 15) >> 4);
		memcpy(out + w, run, 32);
		if (w + 8 != ix->bwt_store.size()) { mc_set_error("index build: inconsistent bwt size"); return MC_ERR_ARG; }
	}

	lap("bwt blocks");
	ix->pac_store.assign((size_t)(G / 4 + 2), 0);
	parallel((G + 3) / 4, [&](int, int64_t b0, int64_t b1) {
		for (int64_t k = b0; k < b1; k++)
		{
			uint8_t v = 0;
			for (int64_t i = 4 * k; i < std::min<int64_t>(G, 4 * k + 4); i++) v |= (uint8_t)((fwd[i] & 3) << ((~i & 3) << 1));
			ix->pac_store[k] = v;
		}
	});
	lap("pac");
	finish_view(ix, primary, L2, (uint64_t)N, G);
	return MC_OK;
}

bool read_exact(FILE* fp, void* dst, size_t n) { return fread(dst, 1, n, fp) == n; }

} // namespace

extern "C" {

int mc_index_build(const uint8_t* fwd_codes, int64_t genome_size, int32_t n_chrom, const int32_t* chrom_len,
                   const char* const* chrom_name, int32_t n_threads, mc_index** out)
{
	if (!fwd_codes || genome_size <= 0 || !out) { mc_set_error("mc_index_build: bad arguments"); return MC_ERR_ARG; }
	if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
	mc_index* ix = new mc_index();
	int64_t tot = 0;
	for (int i = 0; i < n_chrom; i++)
	{
		ix->chrom_len.push_back(chrom_len[i]); tot += chrom_len[i];
		ix->chrom_name.push_back(chrom_name && chrom_name[i] ? chrom_name[i] : ("chr" + std::to_string(i + 1)));
		ix->chrom_anno.push_back(""); ix->chrom_n_ambs.push_back(0);
	}
	if (n_chrom <= 0) { ix->chrom_len.push_back((int32_t)genome_size); ix->chrom_name.push_back("chr1"); ix->chrom_anno.push_back(""); ix->chrom_n_ambs.push_back(0); tot = genome_size; }
	if (tot != genome_size) { delete ix; mc_set_error("mc_index_build: chromosome lengths do not add up to genome_size"); return MC_ERR_ARG; }
	int rc = (2 * genome_size < (1ll << 32)) ? build_core<uint32_t>(ix, fwd_codes, genome_size, n_threads)
	                                         : build_core<uint64_t>(ix, fwd_codes, genome_size, n_threads);
	if (rc != MC_OK) { delete ix; return rc; }
	*out = ix;
	return MC_OK;
}

int mc_index_build_gpu(const uint8_t* fwd_codes, int64_t genome_size, int32_t n_chrom, const int32_t* chrom_len,
                       const char* const* chrom_name, int32_t device, mc_index** out)
{
#ifdef MC_HOSTEMU
	mc_set_error("no GPU in the developer harness"); return MC_ERR_CUDA;
#else
	if (genome_size <= 0 || 2 * genome_size >= (1ll << 34)) { mc_set_error("mc_index_build_gpu: texts of 2^34 symbols and more are sorted on the host (mc_index_build)"); return MC_ERR_ARG; }
	if (device < 0) { mc_set_error("mc_index_build_gpu: bad device"); return MC_ERR_ARG; }
	g_sort_device = device;
	const int rc = mc_index_build(fwd_codes, genome_size, n_chrom, chrom_len, chrom_name, 0, out);   // host threads for packing and the BWT string
	g_sort_device = -1;
	return rc;
#endif
}

int mc_index_build_fasta(const char* fasta_path, int32_t n_threads, mc_index** out)
{
	FILE* fp = fopen(fasta_path, "rb");
	if (!fp) { mc_set_error("cannot open %s", fasta_path); return MC_ERR_IO; }
	std::vector<uint8_t> codes;
	std::vector<int32_t> lens; std::vector<std::string> names, annos; std::vector<int32_t> nambs;
	std::vector<mc_index::Hole> holes;
	srand48(11); // bntseq.c:161-162
	char* line = NULL; size_t cap = 0; ssize_t n;
	int lasts = 0; int64_t seq_off = 0;
	while ((n = getline(&line, &cap, fp)) != -1)
	{
		while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
		if (line[0] == '>')
		{
			char* p = line + 1; char* q = p; while (*q && *q != ' ' && *q != '\t') q++;
			std::string name(p, q - p); while (*q == ' ' || *q == '\t') q++;
			names.push_back(name); annos.push_back(*q ? std::string(q) : std::string("(null)"));
			lens.push_back(0); nambs.push_back(0); lasts = 0; seq_off = (int64_t)codes.size();
			continue;
		}
		if (lens.empty()) continue;
		for (ssize_t i = 0; i < n; i++)
		{
			int ch = (unsigned char)line[i]; if (ch == ' ' || ch == '\t') continue; // kseq drops non-graph characters
			int c = kNt4[ch];
			if (c >= 4)
			{
				if (lasts == ch) holes.back().len++;
				else { holes.push_back({seq_off + lens.back(), 1, (char)ch}); nambs.back()++; }
				c = (int)(lrand48() & 3);
			}
			lasts = ch;
			codes.push_back((uint8_t)c); lens.back()++;
		}
	}
	free(line); fclose(fp);
	if (codes.empty()) { mc_set_error("%s holds no sequence", fasta_path); return MC_ERR_IO; }
	std::vector<const char*> np; for (auto& s : names) np.push_back(s.c_str());
	int rc = mc_index_build(codes.data(), (int64_t)codes.size(), (int32_t)lens.size(), lens.data(), np.data(), n_threads, out);
	if (rc != MC_OK) return rc;
	(*out)->chrom_anno = annos; (*out)->chrom_n_ambs = nambs; (*out)->holes = holes;
	return MC_OK;
}

// fwrite / fprintf / fclose results are all checked: a full disk must not leave a truncated index behind a MC_OK
struct OutFile {
	FILE* fp; bool ok;
	explicit OutFile(const std::string& path, const char* mode) : fp(fopen(path.c_str(), mode)), ok(fp != nullptr) {}
	void put(const void* p, size_t size, size_t n) { if (ok && n && fwrite(p, size, n, fp) != n) ok = false; }
	void printf(const char* fmt, ...) __attribute__((format(printf, 2, 3)))
	{
		if (!ok) return;
		va_list ap; va_start(ap, fmt); if (vfprintf(fp, fmt, ap) < 0) ok = false; va_end(ap);
	}
	bool close() { if (fp && fclose(fp) != 0) ok = false; fp = nullptr; return ok; }
	~OutFile() { if (fp) fclose(fp); }
};

int mc_index_save(const mc_index* ix, const char* prefix)
{
	if (!ix || !prefix) { mc_set_error("mc_index_save: null argument"); return MC_ERR_ARG; }
	const mc_index_view& v = ix->v;
	if (!v.bwt || !v.sa || !v.pac || v.n_sa == 0) { mc_set_error("mc_index_save: the index holds no data"); return MC_ERR_ARG; }
	std::string p(prefix);
	{
		OutFile f(p + ".bwt", "wb");
		f.put(&v.primary, 8, 1); f.put(v.L2 + 1, 8, 4); f.put(v.bwt, 4, v.bwt_size);
		if (!f.close()) { mc_set_error("cannot write %s.bwt", prefix); return MC_ERR_IO; }
	}
	{
		OutFile f(p + ".sa", "wb");
		uint64_t intv = (uint64_t)v.sa_intv;
		f.put(&v.primary, 8, 1); f.put(v.L2 + 1, 8, 4); f.put(&intv, 8, 1); f.put(&v.seq_len, 8, 1); f.put(v.sa + 1, 8, v.n_sa - 1);
		if (!f.close()) { mc_set_error("cannot write %s.sa", prefix); return MC_ERR_IO; }
	}
	const int64_t G = v.genome_size;
	{
		OutFile f(p + ".pac", "wb");
		f.put(v.pac, 1, (size_t)((G >> 2) + ((G & 3) == 0 ? 0 : 1)));
		unsigned char ct = 0; if (G % 4 == 0) f.put(&ct, 1, 1);
		ct = (unsigned char)(G % 4); f.put(&ct, 1, 1);
		if (!f.close()) { mc_set_error("cannot write %s.pac", prefix); return MC_ERR_IO; }
	}
	{
		OutFile f(p + ".ann", "w");
		f.printf("%lld %d %u\n", (long long)G, v.n_chrom, 11u);
		int64_t off = 0;
		for (int i = 0; i < v.n_chrom; i++)
		{
			f.printf("%d %s", 0, ix->chrom_name[i].c_str());
			if (i < (int)ix->chrom_anno.size() && !ix->chrom_anno[i].empty()) f.printf(" %s\n", ix->chrom_anno[i].c_str()); else f.printf(" (null)\n");
			f.printf("%lld %d %d\n", (long long)off, v.chrom_len[i], i < (int)ix->chrom_n_ambs.size() ? ix->chrom_n_ambs[i] : 0);
			off += v.chrom_len[i];
		}
		if (!f.close()) { mc_set_error("cannot write %s.ann", prefix); return MC_ERR_IO; }
	}
	{
		OutFile f(p + ".amb", "w");
		f.printf("%lld %d %u\n", (long long)G, v.n_chrom, (unsigned)ix->holes.size());
		for (auto& h : ix->holes) f.printf("%lld %d %c\n", (long long)h.offset, h.len, h.amb);
		if (!f.close()) { mc_set_error("cannot write %s.amb", prefix); return MC_ERR_IO; }
	}
	return MC_OK;
}

static int index_load_impl(const char* prefix, mc_index** out);
int mc_index_load(const char* prefix, mc_index** out)
{
	if (!prefix || !out) { mc_set_error("mc_index_load: null argument"); return MC_ERR_ARG; }
	try { return index_load_impl(prefix, out); }          // no exception may cross the C ABI
	catch (const std::exception& e) { mc_set_error("mc_index_load(%s): %s", prefix, e.what()); return MC_ERR_IO; }
}
static int index_load_impl(const char* prefix, mc_index** out)
{
	std::string p(prefix);
	std::unique_ptr<mc_index> holder(new mc_index());
	mc_index* ix = holder.get();
	uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0};
	FILE* fp = fopen((p + ".bwt").c_str(), "rb");
	if (!fp) { mc_set_error("cannot open %s.bwt", prefix); return MC_ERR_IO; }
	fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
	if (sz < 40 + 64) { fclose(fp); mc_set_error("%s.bwt is too short to be an index (%ld bytes)", prefix, sz); return MC_ERR_IO; }
	size_t words = ((size_t)sz - 40) >> 2;
	ix->bwt_store.resize(words);
	bool ok = read_exact(fp, &primary, 8) && read_exact(fp, L2 + 1, 32) && read_exact(fp, ix->bwt_store.data(), words * 4);
	fclose(fp);
	if (!ok) { mc_set_error("%s.bwt is truncated", prefix); return MC_ERR_IO; }
	uint64_t seq_len = L2[4];
	if (seq_len == 0 || (seq_len & 1) || seq_len / 16 > words) { mc_set_error("%s.bwt is inconsistent (text length %llu for %zu words)", prefix, (unsigned long long)seq_len, words); return MC_ERR_IO; }
	fp = fopen((p + ".sa").c_str(), "rb");
	if (!fp) { mc_set_error("cannot open %s.sa", prefix); return MC_ERR_IO; }
	uint64_t hdr[7];
	ok = read_exact(fp, hdr, 56);
	uint64_t intv = hdr[5];
	if (!ok || intv != 32 || hdr[6] != seq_len) { fclose(fp); mc_set_error("%s.sa does not match %s.bwt (sa_intv must be 32)", prefix, prefix); return MC_ERR_IO; }
	uint64_t n_sa = (seq_len + intv) / intv;
	ix->sa_store.assign(n_sa, 0); ix->sa_store[0] = (uint64_t)-1;
	ok = read_exact(fp, ix->sa_store.data() + 1, (n_sa - 1) * 8); fclose(fp);
	if (!ok) { mc_set_error("%s.sa is truncated", prefix); return MC_ERR_IO; }
	fp = fopen((p + ".ann").c_str(), "r");
	if (!fp) { mc_set_error("cannot open %s.ann", prefix); return MC_ERR_IO; }
	long long l_pac = 0; int n_seqs = 0; unsigned seed = 0;
	if (fscanf(fp, "%lld%d%u", &l_pac, &n_seqs, &seed) != 3) { fclose(fp); mc_set_error("%s.ann is malformed", prefix); return MC_ERR_IO; }
	for (int i = 0; i < n_seqs; i++)
	{
		unsigned gi; char name[1024]; long long off; int len, nambs;
		if (fscanf(fp, "%u%1023s", &gi, name) != 2) break;
		std::string anno; int c; while ((c = fgetc(fp)) != '\n' && c != EOF) anno.push_back((char)c);
		if (!anno.empty() && anno[0] == ' ') anno.erase(0, 1);
		if (fscanf(fp, "%lld%d%d", &off, &len, &nambs) != 3) break;
		ix->chrom_name.push_back(name); ix->chrom_anno.push_back(anno); ix->chrom_len.push_back(len); ix->chrom_n_ambs.push_back(nambs);
	}
	fclose(fp);
	if ((int)ix->chrom_len.size() != n_seqs || (uint64_t)l_pac * 2 != seq_len) { mc_set_error("%s.ann does not match %s.bwt", prefix, prefix); return MC_ERR_IO; }
	fp = fopen((p + ".amb").c_str(), "r");
	if (fp)
	{
		long long lp; int ns, nh;
		if (fscanf(fp, "%lld%d%d", &lp, &ns, &nh) == 3)
			for (int i = 0; i < nh; i++) { long long off; int len; char s[16]; if (fscanf(fp, "%lld%d%15s", &off, &len, s) != 3) break; ix->holes.push_back({off, len, s[0]}); }
		fclose(fp);
	}
	fp = fopen((p + ".pac").c_str(), "rb");
	if (!fp) { mc_set_error("cannot open %s.pac", prefix); return MC_ERR_IO; }
	ix->pac_store.assign((size_t)(l_pac / 4 + 2), 0);
	size_t got = fread(ix->pac_store.data(), 1, (size_t)(l_pac / 4 + 1), fp); fclose(fp);
	if (got < (size_t)((l_pac + 3) / 4)) { mc_set_error("%s.pac is truncated", prefix); return MC_ERR_IO; }
	finish_view(ix, primary, L2, seq_len, l_pac);
	*out = holder.release();
	return MC_OK;
}

int mc_index_wrap(const mc_index_view* view, mc_index** out)
{
	if (!view || !out || !view->bwt || !view->sa || !view->pac || view->sa_intv != 32) { mc_set_error("mc_index_wrap: bad view (sa_intv must be 32)"); return MC_ERR_ARG; }
	mc_index* ix = new mc_index();
	ix->v = *view;
	for (int i = 0; i < view->n_chrom; i++)
	{
		ix->chrom_len.push_back(view->chrom_len[i]);
		ix->chrom_name.push_back(view->chrom_name && view->chrom_name[i] ? view->chrom_name[i] : ("chr" + std::to_string(i + 1)));
	}
	for (auto& s : ix->chrom_name) ix->chrom_name_ptr.push_back(s.c_str());
	ix->v.chrom_len = ix->chrom_len.data(); ix->v.chrom_name = ix->chrom_name_ptr.data();
	*out = ix;
	return MC_OK;
}

int mc_index_get(const mc_index* idx, mc_index_view* view)
{
	if (!idx || !view) { mc_set_error("mc_index_get: null argument"); return MC_ERR_ARG; }
	*view = idx->v;
	return MC_OK;
}

void mc_index_free(mc_index* idx) { delete idx; }

} // extern "C"
