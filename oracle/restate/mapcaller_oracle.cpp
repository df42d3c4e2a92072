// ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Plain sequential C++ restatement of MapCaller's read-mapping hot path, used by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER of the CUDA path.  Nothing under
// mapcaller_b200/ includes, links or calls this file.  Every function cites the reference code it follows
// (paths relative to /root/reference).  Parity status: PINNED -- tests/test_oracle.py checks this
// restatement record-for-record against the unmodified reference compiled into oracle/_ref (when present)
// and against the committed fixtures under tests/golden/ that were generated from it.
//
// It deliberately works the way the reference does -- one read pair after the other, std::vector /
// std::map containers, avgDist fed back after every 200-read chunk -- and shares no code with the kernels.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

typedef int64_t i64;
typedef uint64_t u64;

// ---------------------------------------------------------------------------------------------------------
// index image (src/structure.h:32-42, src/bwt_index.cpp:105-124,16-36,232-258)
// ---------------------------------------------------------------------------------------------------------
struct Index {
	std::vector<uint32_t> bwt; std::vector<u64> sa; std::vector<uint8_t> pac;
	u64 primary, L2[5], seq_len; i64 G, G2;
	std::vector<i64> chrom_end; std::vector<int> chrom_id;   // PosChrIdMap
	std::string ref;                                          // RefSequence (forward + reverse complement)
};

int nt4(unsigned char c)   // nst_nt4_table, src/BWT_Index/bntseq.c:40-57
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}
char comp(char c)         // GetComplementaryBase, src/tools.cpp:3-17
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
}
void revcomp_inplace(std::string& s)   // SelfComplementarySeq, src/tools.cpp:31-43
{
	std::reverse(s.begin(), s.end());
	for (size_t i = 0; i < s.size(); i++) s[i] = comp(s[i]);
}

bool load_index(const char* prefix, Index& ix)
{
	std::string p(prefix);
	FILE* fp = fopen((p + ".bwt").c_str(), "rb"); if (!fp) return false;
	fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
	ix.L2[0] = 0;
	if (fread(&ix.primary, 8, 1, fp) != 1 || fread(ix.L2 + 1, 8, 4, fp) != 4) { fclose(fp); return false; }
	ix.bwt.resize((sz - 40) / 4);
	if (fread(ix.bwt.data(), 4, ix.bwt.size(), fp) != ix.bwt.size()) { fclose(fp); return false; }
	fclose(fp);
	ix.seq_len = ix.L2[4];
	fp = fopen((p + ".sa").c_str(), "rb"); if (!fp) return false;
	u64 hdr[7]; if (fread(hdr, 8, 7, fp) != 7) { fclose(fp); return false; }
	ix.sa.assign((ix.seq_len + 32) / 32, 0); ix.sa[0] = (u64)-1;
	if (fread(ix.sa.data() + 1, 8, ix.sa.size() - 1, fp) != ix.sa.size() - 1) { fclose(fp); return false; }
	fclose(fp);
	fp = fopen((p + ".ann").c_str(), "r"); if (!fp) return false;
	long long lp; int ns; unsigned seed;
	if (fscanf(fp, "%lld%d%u", &lp, &ns, &seed) != 3) { fclose(fp); return false; }
	ix.G = lp; ix.G2 = 2 * lp;
	std::vector<std::pair<i64, int> > ends; i64 tot = 0;
	for (int i = 0; i < ns; i++)
	{
		unsigned gi; char name[1024]; long long off; int len, na;
		if (fscanf(fp, "%u%1023s", &gi, name) != 2) break;
		int c; while ((c = fgetc(fp)) != '\n' && c != EOF) {}
		if (fscanf(fp, "%lld%d%d", &off, &len, &na) != 3) break;
		i64 fwd = tot; tot += len; i64 rev = ix.G2 - tot;
		ends.push_back(std::make_pair(fwd + len - 1, i)); ends.push_back(std::make_pair(rev + len - 1, i));
	}
	fclose(fp);
	std::sort(ends.begin(), ends.end());
	for (size_t i = 0; i < ends.size(); i++) { ix.chrom_end.push_back(ends[i].first); ix.chrom_id.push_back(ends[i].second); }
	fp = fopen((p + ".pac").c_str(), "rb"); if (!fp) return false;
	ix.pac.assign(ix.G / 4 + 2, 0);
	size_t got = fread(ix.pac.data(), 1, ix.G / 4 + 1, fp); (void)got; fclose(fp);
	ix.ref.resize(ix.G2);
	for (i64 f = 0, r = ix.G2 - 1; f < ix.G; f++, r--)   // IdvLoadReferenceSequences, src/bwt_index.cpp:196-215
	{
		int b = ix.pac[f >> 2] >> ((~f & 3) << 1) & 3;
		ix.ref[f] = "ACGT"[b]; ix.ref[r] = "TGCA"[b];
	}
	return true;
}

// PosChrIdMap.lower_bound
int chrom_lb(const Index& ix, i64 g) { return (int)(std::lower_bound(ix.chrom_end.begin(), ix.chrom_end.end(), g) - ix.chrom_end.begin()); }

// ---------------------------------------------------------------------------------------------------------
// FM-index queries (src/bwt_search.cpp:25-164)
// ---------------------------------------------------------------------------------------------------------
struct Counters { i64 seed_blocks, locate_blocks, sa_reads, dp_cells, dp_tasks; } g_cnt;

// number of symbols equal to c among BWT[0..k] ($ removed), by walking the 128-symbol block of k
u64 occ_upto(const Index& ix, u64 k, int c)
{
	const uint32_t* blk = ix.bwt.data() + ((k >> 7) << 4);
	u64 n = ((const u64*)blk)[c];
	const uint32_t* w = blk + 8;
	for (u64 i = (k >> 7) << 7; i <= k; i++)
	{
		uint32_t word = w[(i & 127) >> 4];
		if ((int)((word >> ((~i & 15) << 1)) & 3) == c) n++;
	}
	return n;
}
void occ4(const Index& ix, u64 k, u64 out[4])   // bwt_occ4, :49-66
{
	if (k == (u64)-1) { out[0] = out[1] = out[2] = out[3] = 0; return; }
	k -= (k >= ix.primary);
	for (int c = 0; c < 4; c++) out[c] = occ_upto(ix, k, c);
}
u64 lf(const Index& ix, u64 k)                  // bwt_invPsi, :101-107
{
	if (k == ix.primary) return 0;
	u64 x = k - (k > ix.primary);
	int c = (ix.bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)] >> ((~x & 15) << 1)) & 3;
	return ix.L2[c] + occ_upto(ix, x, c);
}
u64 locate(const Index& ix, u64 k)              // bwt_sa, :109-119
{
	u64 steps = 0;
	while (k & 31) { steps++; k = lf(ix, k); g_cnt.locate_blocks++; }
	g_cnt.sa_reads++;
	return steps + ix.sa[k >> 5];
}
struct SearchResult { int len, freq; std::vector<u64> loc; };
SearchResult bwt_search(const Index& ix, const std::vector<uint8_t>& code, int start, int stop)   // BWT_Search, :121-164
{
	SearchResult r; r.len = 0; r.freq = 0;
	int p = code[start];
	u64 x0 = ix.L2[p] + 1, x1 = ix.L2[3 - p] + 1, x2 = ix.L2[p + 1] - ix.L2[p];
	int pos = start + 1;
	for (; pos < stop; pos++)
	{
		if (code[pos] > 3) break;
		u64 k = x1 - 1, l = x1 - 1 + x2, tk[4], tl[4];
		occ4(ix, k, tk); occ4(ix, l, tl);
		{   // blocks the reference touches: one when k and l fall into the same 128-row block (:73)
			u64 kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
			g_cnt.seed_blocks += ((kk >> 7) != (ll >> 7) || k == (u64)-1 || l == (u64)-1) ? 2 : 1;
		}
		u64 c0[4], c1[4], c2[4];
		for (int i = 0; i < 4; i++) { c1[i] = ix.L2[i] + 1 + tk[i]; c2[i] = tl[i] - tk[i]; }
		c0[3] = x0 + ((x1 <= ix.primary && x1 + x2 - 1 >= ix.primary) ? 1 : 0);
		c0[2] = c0[3] + c2[3]; c0[1] = c0[2] + c2[2]; c0[0] = c0[1] + c2[1];
		int i = 3 - code[pos];
		if (c2[i] == 0) break;
		x0 = c0[i]; x1 = c1[i]; x2 = c2[i];
	}
	r.len = pos - start;
	if (r.len >= 16 && (int)x2 <= 50) { r.freq = (int)x2; for (int i = 0; i < r.freq; i++) r.loc.push_back(locate(ix, x0 + i)); }
	return r;
}

// ---------------------------------------------------------------------------------------------------------
// per-read structures (src/structure.h:113-150)
// ---------------------------------------------------------------------------------------------------------
struct Frag { bool simple; int rPos; i64 gPos; int rLen, gLen; i64 PosDiff; std::string a1, a2; };
struct Cand { int score; bool orientation; int paired; std::vector<Frag> f; };
struct Read { int rlen; std::string seq; int score, sub_score, best; std::vector<Cand> c; };

struct Params { int max_pos_diff, max_clip, max_dup; float maxmm; bool nw; };

struct State {
	Index ix; Params prm;
	uint32_t avgDist; i64 nReads, nMapped, nPaired, distSum, lenSum;
	std::vector<uint16_t> A, C, G, T, multi, F1, R2, F2, R1; std::vector<uint8_t> rc;
	std::map<i64, std::map<std::string, uint16_t> > ins, del; std::map<i64, uint16_t> bp;
	std::vector<std::pair<i64, i64> > inv, tnl; i64 stale_gpos, stale_dist;
};

// ---------------------------------------------------------------------------------------------------------
// seeding and clustering (src/ReadMapping.cpp:125-226)
// ---------------------------------------------------------------------------------------------------------
bool by_posdiff(const Frag& x, const Frag& y) { return x.PosDiff == y.PosDiff ? x.rPos < y.rPos : x.PosDiff < y.PosDiff; }
bool by_readpos(const Frag& x, const Frag& y) { return x.rPos == y.rPos ? x.gPos < y.gPos : x.rPos < y.rPos; }

std::vector<Frag> simple_pairs(const State& S, const std::string& seq)   // IdentifySimplePairs, :125-158
{
	const int rlen = (int)seq.size();
	std::vector<uint8_t> code(rlen ? rlen : 1);
	for (int i = 0; i < rlen; i++) code[i] = (uint8_t)nt4(seq[i]);
	std::vector<Frag> v;
	int pos = 0;
	while (pos < rlen - 16)
	{
		if (code[pos] > 3) { pos++; continue; }
		SearchResult r = bwt_search(S.ix, code, pos, rlen);
		for (int i = 0; i < r.freq; i++)
		{
			Frag f; f.simple = true; f.rPos = pos; f.gPos = (i64)r.loc[i]; f.rLen = f.gLen = r.len; f.PosDiff = f.gPos - pos;
			if (f.PosDiff > 0) v.push_back(f);
		}
		pos += r.len + 1;
	}
	std::sort(v.begin(), v.end(), by_posdiff);
	return v;   // the reference appends a terminal pair at 2G; the clustering below treats "past the end" the same way
}

std::vector<Cand> cluster(const State& S, int rlen, const std::vector<Frag>& v)   // SimplePairClustering, :194-226
{
	std::vector<Cand> out;
	const int n = (int)v.size();
	if (n == 0) return out;
	int head = 0, score = v[0].rLen, thr = rlen >> 2;
	i64 bound = S.ix.chrom_end[chrom_lb(S.ix, v[0].gPos)];
	for (int j = 1; j <= n; j++)
	{
		bool cut = (j == n) || v[j].gPos > bound || std::llabs(v[j].PosDiff - v[j - 1].PosDiff) > S.prm.max_pos_diff;
		if (!cut) { score += v[j].rLen; continue; }
		if (score > thr)
		{
			if (thr < (score >> 1)) thr = score >> 1;
			Cand c; c.paired = -1; c.orientation = true;
			if (score >= rlen)   // tandem repeat: IdentifyClosestFragmentPairs, :160-192
			{
				int bi = head, bj = head, best = 0, ri = head, s = v[head].rLen;
				for (int k = head + 1; k < j; k++)
				{
					if (v[k].PosDiff != v[ri].PosDiff) { if (s > best) { best = s; bi = ri; bj = k; } ri = k; s = v[k].rLen; }
					else s += v[k].rLen;
				}
				if (s > best) { best = s; bi = ri; bj = j; }
				c.score = best; c.f.assign(v.begin() + bi, v.begin() + bj);
			}
			else { c.score = score; c.f.assign(v.begin() + head, v.begin() + j); }
			out.push_back(c);
		}
		if (j < n) { head = j; score = v[j].rLen; bound = S.ix.chrom_end[chrom_lb(S.ix, v[j].gPos)]; }
	}
	return out;
}

// ---------------------------------------------------------------------------------------------------------
// pairing (src/ReadMapping.cpp:228-322)
// ---------------------------------------------------------------------------------------------------------
void keep_best_only(std::vector<Cand>& v)   // RemoveRedundantAlnCan
{
	if (v.size() <= 1) return;
	int mx = 0; for (auto& c : v) mx = std::max(mx, c.score);
	for (auto& c : v) if (c.score < mx) c.score = 0;
}
int pair_up(i64 est, std::vector<Cand>& a, std::vector<Cand>& b)   // CheckPairedAlignmentDistance
{
	if (a.size() * b.size() > 100) { keep_best_only(a); keep_best_only(b); }
	std::vector<int> partner(a.size(), -1); i64 top = 0;
	for (size_t i = 0; i < a.size(); i++)
	{
		if (a[i].score == 0) continue;
		int best = -1, bs = 0;
		for (size_t j = 0; j < b.size(); j++)
		{
			if (b[j].score == 0 || b[j].f[0].PosDiff < a[i].f[0].PosDiff) continue;
			if (b[j].f[0].PosDiff - a[i].f[0].PosDiff < est && b[j].score > bs) { best = (int)j; bs = b[j].score; }
		}
		partner[i] = best;
		if (best >= 0) top = std::max<i64>(top, a[i].score + b[best].score);
	}
	int n = 0;
	if (top > 0)
		for (size_t i = 0; i < a.size(); i++)
			if (partner[i] >= 0 && a[i].score + b[partner[i]].score == top) { n++; a[i].paired = partner[i]; b[partner[i]].paired = (int)i; }
	return n;
}
void mask_unpaired(std::vector<Cand>& a, std::vector<Cand>& b)   // MaskUnPairedAlnCan
{
	int top = 0;
	for (auto& c : a) if (c.paired != -1) top = std::max(top, c.score + b[c.paired].score);
	for (auto& c : a) if (c.paired == -1 || c.score + b[c.paired].score < top) c.score = 0;
	for (auto& c : b) if (c.paired == -1 || c.score + a[c.paired].score < top) c.score = 0;
}

// ---------------------------------------------------------------------------------------------------------
// rescue (src/AlignmentRescue.cpp, src/KmerAnalysis.cpp)
// ---------------------------------------------------------------------------------------------------------
struct Kmer { uint32_t wid, pos; };
uint32_t kmer_id(const char* s, int p) { uint32_t id = 0; for (int i = p; i < p + 8; i++) id = (id << 2) + nt4(s[i]); return id; }
std::vector<Kmer> kmers_of(const char* s, int len)   // CreateKmerVecFromReadSeq, src/KmerAnalysis.cpp:57-103 (quirk after 'N' included)
{
	std::vector<Kmer> v;
	uint32_t count = 0, head, tail = 0;
	while (count < 8 && tail < (uint32_t)len) { if (s[tail++] != 'N') count++; else count = 0; }
	if (count != 8) return v;
	Kmer k; k.pos = (head = tail - 8); k.wid = kmer_id(s, head); v.push_back(k);
	for (head += 1; tail < (uint32_t)len; head++, tail++)
	{
		if (s[tail] != 'N') { k.pos = head; k.wid = ((k.wid & 0x3FFF) << 2) + nt4(s[tail]); v.push_back(k); }
		else
		{
			count = 0; tail++;
			while (count < 8 && tail < (uint32_t)len) { if (s[tail++] != 'N') count++; else count = 0; }
			if (count != 8) break;
			k.pos = (head = tail - 8); k.wid = kmer_id(s, head); v.push_back(k);
		}
	}
	std::stable_sort(v.begin(), v.end(), [](const Kmer& x, const Kmer& y) { return x.wid < y.wid; });
	return v;
}
struct KPair { int diff; uint32_t r, g; };
bool rescue_window(State& S, const std::vector<Kmer>& rk, int rlen, i64 left, i64 right, Cand& out)
{
	if (right > S.ix.G2) right = S.ix.G2;
	int i1 = chrom_lb(S.ix, left), i2 = chrom_lb(S.ix, right);
	// right == 2G: the reference reads PosChrIdMap.end()->second (src/AlignmentRescue.cpp:62-63), which with GCC's layout of the
	// globals of src/main.cpp is the zeroed first word of the next map, i.e. chromosome id 0 (pinned against oracle/_ref)
	if (i1 >= (int)S.ix.chrom_end.size()) return false;
	if (S.ix.chrom_id[i1] != (i2 >= (int)S.ix.chrom_end.size() ? 0 : S.ix.chrom_id[i2])) return false;
	const i64 slen = right - left;
	if (slen < rlen) return false;
	// window k-mers; positions before the start of RefSequence (left < 0) hold no usable text
	std::string win((size_t)slen, 'N');
	for (i64 i = 0; i < slen; i++) if (left + i >= 0) win[i] = S.ix.ref[left + i];
	std::vector<Kmer> wk = kmers_of(win.data(), (int)slen);
	std::vector<KPair> kp;   // IdentifyCommonKmers, :105-131
	for (auto& a : rk)
	{
		auto it = std::lower_bound(wk.begin(), wk.end(), a, [](const Kmer& x, const Kmer& y) { return x.wid < y.wid; });
		for (; it != wk.end() && it->wid == a.wid; ++it)
		{
			i64 d = (i64)it->pos - (i64)a.pos;
			if (std::llabs(d) < slen) { KPair p; p.r = a.pos; p.g = it->pos; p.diff = (int)d; kp.push_back(p); }
		}
	}
	std::sort(kp.begin(), kp.end(), [](const KPair& x, const KPair& y) { return x.diff == y.diff ? x.r < y.r : x.diff < y.diff; });
	std::vector<Frag> seeds;   // GenerateSimplePairsFromCommonKmers, :133-163
	for (size_t i = 0; i < kp.size();)
	{
		size_t j = i + 1; uint32_t next = kp[i].r + 1;
		while (j < kp.size() && kp[j].r == next && kp[j].diff == kp[i].diff) { j++; next++; }
		int l = 8 + (int)(j - 1 - i);
		if (l >= 10) { Frag f; f.simple = true; f.rPos = (int)kp[i].r; f.gPos = kp[i].g + left; f.PosDiff = kp[i].diff + left; f.rLen = f.gLen = l; seeds.push_back(f); }
		i = j;
	}
	if (seeds.empty()) return false;
	out.score = 0; out.f.clear();   // IdentifyBestAlnCan, src/AlignmentRescue.cpp:3-26
	for (size_t i = 0; i < seeds.size();)
	{
		size_t j = i + 1; int s = seeds[i].rLen;
		while (j < seeds.size() && seeds[j].PosDiff == seeds[i].PosDiff) { s += seeds[j].rLen; j++; }
		if (s > out.score) { out.score = s; out.f.assign(seeds.begin() + i, seeds.begin() + j); }
		i = j;
	}
	return true;
}
int rescue(State& S, uint32_t est, Read& r1, Read& r2)   // AlignmentRescue, src/AlignmentRescue.cpp:28-111
{
	int s1 = 0, s2 = 0, n = 0;
	for (auto& c : r1.c) s1 = std::max(s1, c.score);
	for (auto& c : r2.c) s2 = std::max(s2, c.score);
	if (s1 < (r1.rlen >> 2) && s2 < (r2.rlen >> 2)) return 0;
	int strat = (s1 - s2 > (r2.rlen >> 2)) ? 1 : (s2 - s1 > (r1.rlen >> 2)) ? 2 : 3;
	int num1 = (int)r1.c.size(), num2 = (int)r2.c.size();
	if (strat == 1 || strat == 3)
	{
		std::vector<Kmer> rk = kmers_of(r2.seq.data(), r2.rlen);
		for (size_t i = 0; i < r1.c.size(); i++)
		{
			if (r1.c[i].score < (s1 >> 1) || r1.c[i].paired != -1) continue;
			i64 d = r1.c[i].f[0].PosDiff; Cand c;
			if (!rescue_window(S, rk, r2.rlen, d, d + est + r2.rlen, c)) continue;
			if (c.score > s2) { n++; r1.c[i].paired = num2++; c.paired = (int)i; c.orientation = true; r2.c.push_back(c); }
		}
	}
	if (strat == 2 || strat == 3)
	{
		std::vector<Kmer> rk = kmers_of(r1.seq.data(), r1.rlen);
		for (size_t j = 0; j < r2.c.size(); j++)
		{
			if (r2.c[j].score < (s2 >> 1) || r2.c[j].paired != -1) continue;
			i64 d = r2.c[j].f[0].PosDiff; Cand c;
			if (!rescue_window(S, rk, r1.rlen, d - (i64)est, d + r1.rlen, c)) continue;
			if (c.score > s1) { n++; r2.c[j].paired = num1++; c.paired = (int)j; c.orientation = true; r1.c.push_back(c); }
		}
	}
	return n;
}

// ---------------------------------------------------------------------------------------------------------
// gapped fills
// ---------------------------------------------------------------------------------------------------------
// nw_alignment (src/nw_alignment.cpp:18-83) in exact doubled integers: +-2 (mis)match, gap of length k costs 2+k
void nw_align(std::string& s1, std::string& s2)
{
	const int m = (int)s1.size(), n = (int)s2.size();
	std::vector<std::vector<int> > r(m + 1, std::vector<int>(n + 1)), t = r, s = r;
	const int NEG = -131072;
	for (int i = 1; i <= m; i++) { r[i][0] = NEG; s[i][0] = t[i][0] = -2 - i; }
	for (int j = 1; j <= n; j++) { t[0][j] = NEG; s[0][j] = r[0][j] = -2 - j; }
	for (int i = 1; i <= m; i++)
		for (int j = 1; j <= n; j++)
		{
			r[i][j] = std::max(r[i][j - 1] - 1, s[i][j - 1] - 3);
			t[i][j] = std::max(t[i - 1][j] - 1, s[i - 1][j] - 3);
			int d = s[i - 1][j - 1] + (nt4(s1[i - 1]) == nt4(s2[j - 1]) ? 2 : -2);
			s[i][j] = std::max(d, std::max(r[i][j], t[i][j]));
		}
	int i = m, j = n;
	while (i > 0 || j > 0)
	{
		if (s[i][j] == r[i][j]) { s1.insert(i, 1, '-'); j--; }
		else if (s[i][j] == t[i][j]) { s2.insert(j, 1, '-'); i--; }
		else { i--; j--; }
	}
	g_cnt.dp_cells += (i64)m * n; g_cnt.dp_tasks++;
}
// ksw2_alignment (src/ksw2_alignment.cpp:250-272) = ksw_extz2_sse with full band + ksw_backtrack, restated as the
// plain affine recurrence it vectorises
void ksw2_align(std::string& s1, std::string& s2)
{
	const int qlen = (int)s1.size(), tlen = (int)s2.size();   // query = read piece, target = genome piece
	const int NEG = -0x20000000;
	std::vector<int> H(qlen), E(qlen, NEG);
	std::vector<uint8_t> dir((size_t)qlen * tlen);
	for (int j = 0; j < qlen; j++) H[j] = -(2 + (j + 1));
	for (int i = 0; i < tlen; i++)
	{
		int diag = i == 0 ? 0 : -(2 + i), left = -(2 + (i + 1)), F = NEG;
		const int ct = nt4(s2[i]);
		for (int j = 0; j < qlen; j++)
		{
			const int cq = nt4(s1[j]);
			const int sc = (ct == 4 || cq == 4) ? 0 : (ct == cq ? 1 : -1);
			int e = i > 0 ? std::max(H[j] - 2, E[j]) - 1 : NEG;
			int f = j > 0 ? std::max(left - 2, F) - 1 : NEG;
			int z = diag + sc; uint8_t d = 0;
			if (e > z) { d = 1; z = e; }
			if (f > z) { d = 2; z = f; }
			if (e > z - 2) d |= 0x08;
			if (f > z - 2) d |= 0x10;
			dir[(size_t)i * qlen + j] = d;
			diag = H[j]; H[j] = z; E[j] = e; left = z; F = f;
		}
	}
	std::string cigar;   // ksw_backtrack, :25-68
	int i = tlen - 1, j = qlen - 1, state = 0;
	while (i >= 0 && j >= 0)
	{
		uint32_t tmp = dir[(size_t)i * qlen + j];
		if (state == 0) state = tmp & 7; else if (!((tmp >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (state == 0) { cigar.push_back('M'); i--; j--; }
		else if (state == 1 || state == 3) { cigar.push_back('D'); i--; }
		else { cigar.push_back('I'); j--; }
	}
	if (i >= 0) cigar.append((size_t)(i + 1), 'D');
	if (j >= 0) cigar.append((size_t)(j + 1), 'I');
	int p = 0;
	for (int k = (int)cigar.size() - 1; k >= 0; k--, p++)
	{
		if (cigar[k] == 'D') s1.insert(s1.begin() + p, '-');
		else if (cigar[k] == 'I') s2.insert(s2.begin() + p, '-');
	}
	g_cnt.dp_cells += (i64)qlen * tlen; g_cnt.dp_tasks++;
}

// ---------------------------------------------------------------------------------------------------------
// candidate -> alignment (src/ReadAlignment.cpp)
// ---------------------------------------------------------------------------------------------------------
void fill_piece(State& S, const std::string& seq, Frag& f)   // ProcessNormalPair, :155-191
{
	f.a1 = f.rLen > 0 ? seq.substr(f.rPos, f.rLen) : std::string((size_t)f.gLen, '-');
	f.a2 = f.gLen > 0 ? S.ix.ref.substr(f.gPos, f.gLen) : std::string((size_t)f.rLen, '-');
	if (f.gPos >= S.ix.G) { if (f.rLen > 0) revcomp_inplace(f.a1); if (f.gLen > 0) revcomp_inplace(f.a2); }
	if (f.rLen > 0 && f.gLen > 0)
	{
		bool dp = f.rLen != f.gLen;
		if (!dp) { int mis = 0; for (int i = 0; i < f.rLen; i++) if (f.a1[i] != f.a2[i]) mis++; dp = mis > 1 && mis >= (int)(f.rLen * 0.2); }
		if (dp) { if (S.prm.nw) nw_align(f.a1, f.a2); else ksw2_align(f.a1, f.a2); }
	}
}
void trim(Frag& f, bool heading, bool first)   // RemoveHeadingGaps / RemoveTailingGaps, :264-304
{
	int rs = 0, gs = 0, cut = 0, len = (int)f.a1.size();
	for (int k = 0; k < len; k++)
	{
		int j = heading ? k : len - 1 - k;
		if (f.a1[j] == '-') gs++; else if (f.a2[j] == '-') rs++; else break;
		cut++;
	}
	if (!cut) return;
	if (heading) { f.a1.erase(0, cut); f.a2.erase(0, cut); } else { f.a1.resize(len - cut); f.a2.resize(len - cut); }
	f.rLen -= rs; f.gLen -= gs;
	if (first) { f.rPos += rs; f.gPos += gs; }
}
bool piece_ok(const Frag& f)   // CheckLocalAlignmentQuality, :193-232
{
	int kind = -1, n = 0, mis = 0, changes = 0;
	for (size_t i = 0; i < f.a1.size(); i++)
	{
		int k = f.a1[i] == '-' ? 0 : f.a2[i] == '-' ? 1 : 2;
		if (k == 2) { n++; if (f.a1[i] != f.a2[i]) mis++; }
		if (k != kind) { kind = k; changes++; }
	}
	return !(changes >= 4 || (mis >= 3 && mis >= (int)(n * 0.3)));
}
bool produce_alignment(State& S, Read& rd)   // ProduceReadAlignment, :306-430
{
	const int max_mm = (int)(rd.rlen * S.prm.maxmm);
	for (size_t ci = 0; ci < rd.c.size(); ci++)
	{
		Cand& c = rd.c[ci];
		if (c.score == 0) continue;
		std::vector<Frag>& v = c.f;
		std::sort(v.begin(), v.end(), by_readpos);
		bool ov = false;   // RemoveOverlaps, :38-65
		for (size_t i = 0; i + 1 < v.size(); i++)
		{
			if (v[i].rPos == v[i + 1].rPos) { ov = true; v[i].rLen = v[i].gLen = 0; }
			else if (v[i].gPos >= v[i + 1].gPos || v[i].gPos + v[i].gLen > v[i + 1].gPos)
			{
				ov = true; int o = (int)(v[i].gPos + v[i].gLen - v[i + 1].gPos);
				v[i].rLen = std::max(0, v[i].rLen - o); v[i].gLen = std::max(0, v[i].gLen - o);
			}
		}
		if (ov) v.erase(std::remove_if(v.begin(), v.end(), [](const Frag& f) { return f.rLen == 0; }), v.end());
		{   // IdentifyNormalPairs, :67-108
			std::vector<Frag> gaps; const size_t n0 = v.size();
			for (size_t i = 0; i + 1 < n0; i++)
			{
				int rg = std::max(0, v[i + 1].rPos - (v[i].rPos + v[i].rLen));
				int gg = (int)std::max<i64>(0, v[i + 1].gPos - (v[i].gPos + v[i].gLen));
				if (rg > 0 || gg > 0) { Frag f; f.simple = false; f.rPos = v[i].rPos + v[i].rLen; f.gPos = v[i].gPos + v[i].gLen; f.PosDiff = f.gPos - f.rPos; f.rLen = rg; f.gLen = gg; gaps.push_back(f); }
			}
			if (!gaps.empty()) { v.insert(v.end(), gaps.begin(), gaps.end()); std::inplace_merge(v.begin(), v.begin() + n0, v.end(), by_readpos); }
			if (v[0].rPos > 0) { Frag f; f.simple = false; f.rPos = 0; f.gPos = f.PosDiff = v[0].PosDiff; f.rLen = f.gLen = v[0].rPos; v.insert(v.begin(), f); }
			const Frag& last = v.back();
			if (last.rPos + last.rLen < rd.rlen) { Frag f; f.simple = false; f.rPos = last.rPos + last.rLen; f.gPos = last.gPos + last.gLen; f.PosDiff = last.PosDiff; f.rLen = f.gLen = rd.rlen - f.rPos; v.push_back(f); }
		}
		{   // CheckAlignmentValidity, src/tools.cpp:119-130
			i64 g0 = v.front().gPos, g1 = v.back().gPos + v.back().gLen; bool ok = !(g0 < 0 || g1 > S.ix.G2);
			if (ok) { int a = chrom_lb(S.ix, g0), b = chrom_lb(S.ix, g1 - 1); ok = a < (int)S.ix.chrom_end.size() && b < (int)S.ix.chrom_end.size() && S.ix.chrom_end[a] == S.ix.chrom_end[b]; }
			if (!ok) { c.score = 0; continue; }
		}
		bool head = true, tail = true; const int nf = (int)v.size();
		for (int i = 0; i < nf; i++)
		{
			if (v[i].simple) continue;
			fill_piece(S, rd.seq, v[i]);
			if (i == 0)
			{
				trim(v[i], v[i].gPos < S.ix.G, true);
				if (v[i].a1.size() >= 5 && !piece_ok(v[i])) { head = false; v[i].rLen = v[i].gLen = 0; v[i].a1.clear(); v[i].a2.clear(); v[i].rPos = v[1].rPos; v[i].gPos = v[1].gPos; }
			}
			else if (i == nf - 1)
			{
				trim(v[i], !(v[i].gPos < S.ix.G), false);
				if (v[i].a1.size() >= 5 && !piece_ok(v[i])) { tail = false; v[i].rLen = v[i].gLen = 0; v[i].a1.clear(); v[i].a2.clear(); v[i].rPos = v[i - 1].rPos + v[i - 1].rLen; v[i].gPos = v[i - 1].gPos + v[i - 1].gLen; }
			}
			else if (v[i].rLen >= 5 && v[i].gLen >= 5 && !piece_ok(v[i])) { c.score = 0; break; }
		}
		if (c.score == 0) continue;
		if (!head && !tail) { c.score = 0; continue; }
		int score = 0, mism = 0;   // EvaluateAlignmentScore :234, FindMisMatchNumber :247
		for (auto& f : v)
		{
			if (f.simple) { score += f.rLen; continue; }
			for (size_t k = 0; k < f.a1.size(); k++) { if (f.a1[k] == f.a2[k]) score++; else if (f.a1[k] != '-' && f.a2[k] != '-') mism++; }
		}
		c.score = score;
		if (score == 0) continue;
		if (score < (int)(rd.rlen * (1 - S.prm.maxmm)) && mism > max_mm) { c.score = 0; continue; }
		c.orientation = v[0].gPos < S.ix.G;
		if (!c.orientation) std::reverse(v.begin(), v.end());
		if (c.score > rd.score) { rd.score = c.score; rd.best = (int)ci; }
		else if (c.score > rd.sub_score) rd.sub_score = c.score;
	}
	for (auto& c : rd.c) if (c.score < rd.score) c.score = 0;
	return rd.score > 0;
}

// ---------------------------------------------------------------------------------------------------------
// profile (src/AlignmentProfile.cpp)
// ---------------------------------------------------------------------------------------------------------
void bump12(std::vector<uint16_t>& v, i64 g) { if (g >= 0 && g < (i64)v.size() && v[g] < 4095) v[g]++; }
void count_base(State& S, i64 g, char b)
{
	switch (b) { case 'A': bump12(S.A, g); break; case 'C': bump12(S.C, g); break; case 'G': bump12(S.G, g); break; case 'T': bump12(S.T, g); break; }
}
void update_profile(State& S, bool first_mate, const Read& rd)   // UpdateProfile, :41-242
{
	const i64 G = S.ix.G, G2 = S.ix.G2;
	for (auto& c : rd.c)
	{
		if (c.score == 0) continue;
		const Frag &fb = c.f.front(), &fe = c.f.back();
		if (fb.rLen == 0 && fb.gLen == 0)
		{
			if (fb.rPos > 20) S.bp[fb.gPos < G ? fb.gPos : G2 - 1 - fb.gPos]++;
			if (fb.rPos > S.prm.max_clip) continue;
		}
		if (fe.rLen == 0 && fe.gLen == 0)
		{
			if (rd.rlen - fe.rPos > 20) S.bp[fe.gPos < G ? fe.gPos : G2 - 1 - fe.gPos]++;
			if (rd.rlen - fe.rPos > S.prm.max_clip) continue;
		}
		i64 start = c.orientation ? fb.gPos : G2 - (fb.gPos + fb.gLen);
		if (start < 0 || start >= G) continue;                     // outside the array in the reference
		if (S.rc[start] < S.prm.max_dup) S.rc[start]++; else continue;
		std::vector<uint16_t>& strand = first_mate ? (c.orientation ? S.F1 : S.R1) : (c.orientation ? S.R2 : S.F2);
		for (int i = 0; i < rd.rlen; i++) if (start + i < G) strand[start + i]++;
		for (auto& f : c.f)
		{
			if (f.simple)
			{
				for (int j = 0; j < f.rLen; j++)
				{
					if (c.orientation) count_base(S, f.gPos + j, rd.seq[f.rPos + j]);
					else { char b = rd.seq[f.rPos + j]; if (b == 'A' || b == 'C' || b == 'G' || b == 'T') count_base(S, G2 - 1 - f.gPos - j, comp(b)); }
				}
			}
			else if (f.gLen == 0) S.ins[(c.orientation ? f.gPos : G2 - f.gPos) - 1][f.a1]++;
			else if (f.rLen == 0) S.del[(c.orientation ? f.gPos : G2 - f.gPos - f.gLen) - 1][f.a2]++;
			else
			{
				i64 g = c.orientation ? f.gPos : G2 - (f.gPos + f.gLen);
				for (size_t j = 0; j < f.a1.size();)
				{
					if (f.a2[j] == '-') { size_t e = 1; while (j + e < f.a2.size() && f.a2[j + e] == '-') e++; S.ins[g - 1][f.a1.substr(j, e)]++; j += e; }
					else if (f.a1[j] == '-') { size_t e = 1; while (j + e < f.a1.size() && f.a1[j + e] == '-') e++; S.del[g - 1][f.a2.substr(j, e)]++; j += e; g += e; }
					else { count_base(S, g, f.a1[j]); j++; g++; }
				}
			}
		}
	}
}
void update_multihit(State& S, const Read& rd)   // UpdateMultiHitCount, :244-271
{
	for (auto& c : rd.c)
	{
		if (c.score <= 0) continue;
		i64 g0, g1;
		if (c.orientation) { g0 = c.f.front().gPos; g1 = c.f.back().gPos + c.f.back().gLen; }
		else { g0 = S.ix.G2 - (c.f.front().gPos + c.f.front().gLen); g1 = S.ix.G2 - c.f.back().gPos; }
		for (i64 g = g0; g < g1; g++) bump12(S.multi, g);
	}
}

// ---------------------------------------------------------------------------------------------------------
// the thread body (src/ReadMapping.cpp:416-646)
// ---------------------------------------------------------------------------------------------------------
struct Blob { std::vector<uint8_t> b; template <class T> void put(T v) { const uint8_t* p = (const uint8_t*)&v; b.insert(b.end(), p, p + sizeof(T)); } };
void dump_read(Blob& o, const Read& r)   // same record layout as oracle/ref_shim.cpp:serialise_read
{
	o.put<int32_t>(r.rlen); o.put<int32_t>(r.score); o.put<int32_t>(r.sub_score); o.put<int32_t>(r.best); o.put<int32_t>((int32_t)r.c.size());
	for (auto& c : r.c)
	{
		o.put<int32_t>(c.score); o.put<int32_t>(c.score > 0 ? (c.orientation ? 1 : 0) : -1); o.put<int32_t>(c.paired);
		int32_t nf = c.score > 0 ? (int32_t)c.f.size() : 0; o.put<int32_t>(nf);
		for (int32_t i = 0; i < nf; i++)
		{
			const Frag& f = c.f[i];
			o.put<int32_t>(f.simple ? 1 : 0); o.put<int32_t>(f.rPos); o.put<int64_t>(f.gPos); o.put<int32_t>(f.rLen); o.put<int32_t>(f.gLen);
			int32_t al = f.simple ? 0 : (int32_t)f.a1.size(); o.put<int32_t>(al);
			o.b.insert(o.b.end(), f.a1.begin(), f.a1.begin() + al); o.b.insert(o.b.end(), f.a2.begin(), f.a2.begin() + al);
		}
	}
}
i64 first_gpos(const Cand& c) { return c.f[0].gPos; }

void map_reads(State& S, i64 n, const char* seq, const i64* off, bool paired, bool profile, Blob& out)
{
	std::vector<int32_t> est_log;
	for (i64 base = 0; base < n; base += 200)
	{
		const int cnt = (int)std::min<i64>(200, n - base);
		std::vector<Read> R(cnt);
		for (int i = 0; i < cnt; i++) { R[i].seq.assign(seq + off[base + i], seq + off[base + i + 1]); R[i].rlen = (int)R[i].seq.size(); R[i].score = R[i].sub_score = 0; R[i].best = -1; }
		int mapped = 0, pairedn = 0; i64 dsum = 0, lsum = 0;
		if (paired && cnt % 2 == 0)
		{
			const int est = (int)(S.avgDist * 1.5); est_log.push_back(est);
			for (int i = 0; i < cnt; i += 2)
			{
				Read &a = R[i], &b = R[i + 1];
				a.c = cluster(S, a.rlen, simple_pairs(S, a.seq));
				revcomp_inplace(b.seq);                       // ReverseOrientation, src/tools.cpp:45-55
				b.c = cluster(S, b.rlen, simple_pairs(S, b.seq));
				int np = pair_up(est, a.c, b.c);
				if (np == 0) np = rescue(S, (uint32_t)est, a, b);
				if (np == 0) { keep_best_only(a.c); keep_best_only(b.c); } else mask_unpaired(a.c, b.c);
				if (produce_alignment(S, a)) mapped++;
				if (produce_alignment(S, b)) mapped++;
				// GenCoordinatePair, :343-394
				i64 g1 = 0, g2 = 0, dist = 0;
				for (auto& c : a.c) if (c.score > 0 && c.paired != -1 && b.c[c.paired].score > 0) { g1 = first_gpos(c); g2 = first_gpos(b.c[c.paired]); dist = std::llabs(g2 - g1); break; }
				if (dist == 0)
				{
					std::vector<i64> v1, v2;
					for (auto& c : a.c) if (c.score > 0) v1.push_back(first_gpos(c));
					for (auto& c : b.c) if (c.score > 0) v2.push_back(first_gpos(c));
					if (v1.size() == 1 && v2.size() == 1) { g1 = v1[0]; g2 = v2[0]; dist = std::llabs(g2 - g1); }
					else if (v1.empty() && !v2.empty()) { g1 = -1; dist = g2 = v2[0]; }
					else if (!v1.empty() && v2.empty()) { dist = g1 = v1[0]; g2 = -1; }
				}
				if (dist == 0 || g1 == -1 || g2 == -1) continue;
				const i64 G = S.ix.G, G2 = S.ix.G2;   // classification, :486-532
				if (g1 < G && g2 >= G)
				{
					if (profile) { S.stale_dist = std::llabs(G2 - g1 - g2); if (S.stale_dist > 1000 && S.stale_dist < 10000000) { S.stale_gpos = g1; S.inv.push_back(std::make_pair(S.stale_gpos, S.stale_dist)); } }
				}
				else if (g1 >= G && g2 < G)
				{
					if (profile) { S.stale_dist = std::llabs(G2 - g1 - g2); if (S.stale_dist > 1000 && S.stale_dist < 10000000) S.stale_gpos = g2; S.inv.push_back(std::make_pair(S.stale_gpos, S.stale_dist)); }   // :502 pushes unconditionally
				}
				else if (dist > 1000)
				{
					if (profile)
					{
						S.stale_dist = dist;
						if (g1 < G && g2 < G) { S.tnl.push_back(std::make_pair(g1, dist)); S.tnl.push_back(std::make_pair(g2, dist)); S.stale_gpos = g2; }
						else if (g1 >= G && g2 >= G) { S.tnl.push_back(std::make_pair(G2 - g1, dist)); S.tnl.push_back(std::make_pair(G2 - g2, dist)); S.stale_gpos = G2 - g2; }
					}
				}
				else { lsum += a.rlen + b.rlen; pairedn++; dsum += dist; }
			}
			S.nReads += cnt; S.nMapped += mapped; S.nPaired += pairedn; S.distSum += dsum; S.lenSum += lsum;
			if (S.nPaired > 1000) S.avgDist = (uint32_t)(int)(1. * S.distSum / S.nPaired + .5);
		}
		else
		{
			for (int i = 0; i < cnt; i++) { R[i].c = cluster(S, R[i].rlen, simple_pairs(S, R[i].seq)); keep_best_only(R[i].c); if (produce_alignment(S, R[i])) mapped++; }
			S.nReads += cnt; S.nMapped += mapped;
		}
		if (profile)
			for (int i = 0; i < cnt; i++)
			{
				if (R[i].score == 0) continue;
				int live = 0; for (auto& c : R[i].c) if (c.score > 0) live++;
				if (live == 1) update_profile(S, paired ? (i % 2 == 0) : true, R[i]); else update_multihit(S, R[i]);
			}
		for (int i = 0; i < cnt; i++) dump_read(out, R[i]);
	}
	for (size_t i = 0; i < est_log.size(); i++) out.put<int32_t>(est_log[i]);
}

uint8_t* release(Blob& b, i64* n) { uint8_t* p = (uint8_t*)malloc(b.b.size() ? b.b.size() : 1); memcpy(p, b.b.data(), b.b.size()); *n = (i64)b.b.size(); return p; }

} // namespace

extern "C" {

void* mco_create(const char* prefix, int max_pos_diff, int max_clip, int max_dup, float maxmm, int nw)
{
	State* S = new State();
	if (!load_index(prefix, S->ix)) { delete S; return 0; }
	S->prm.max_pos_diff = max_pos_diff; S->prm.max_clip = max_clip; S->prm.max_dup = max_dup; S->prm.maxmm = maxmm; S->prm.nw = nw != 0;
	S->avgDist = 1000; S->nReads = S->nMapped = S->nPaired = S->distSum = S->lenSum = 0; S->stale_gpos = S->stale_dist = 0;
	const size_t G = (size_t)S->ix.G;
	S->A.assign(G, 0); S->C = S->G = S->T = S->multi = S->F1 = S->R2 = S->F2 = S->R1 = S->A; S->rc.assign(G, 0);
	memset(&g_cnt, 0, sizeof(g_cnt));
	return S;
}
void mco_destroy(void* h) { delete (State*)h; }
int64_t mco_genome_size(void* h) { return ((State*)h)->ix.G; }

uint8_t* mco_map(void* h, int64_t n, const char* seq, const int64_t* off, int paired, int profile, int64_t* nbytes)
{
	Blob out; map_reads(*(State*)h, n, seq, off, paired != 0, profile != 0, out);
	return release(out, nbytes);
}
void mco_counters(void* h, int64_t out[8])
{
	State* S = (State*)h;
	out[0] = S->nReads; out[1] = S->nMapped; out[2] = S->nPaired; out[3] = S->distSum; out[4] = S->lenSum; out[5] = S->avgDist; out[6] = 0; out[7] = 0;
}
void mco_work(int64_t out[5]) { out[0] = g_cnt.seed_blocks; out[1] = g_cnt.locate_blocks; out[2] = g_cnt.sa_reads; out[3] = g_cnt.dp_cells; out[4] = g_cnt.dp_tasks; }
void mco_profile(void* h, int64_t beg, int64_t end, int32_t* out)
{
	State* S = (State*)h;
	for (int64_t g = beg; g < end; g++)
	{
		int32_t* o = out + (g - beg) * 10;
		o[0] = S->A[g]; o[1] = S->C[g]; o[2] = S->G[g]; o[3] = S->T[g]; o[4] = S->multi[g]; o[5] = S->rc[g]; o[6] = S->F1[g]; o[7] = S->R2[g]; o[8] = S->F2[g]; o[9] = S->R1[g];
	}
}
uint8_t* mco_indels(void* h, int which, int64_t* nbytes)
{
	State* S = (State*)h; Blob o;
	for (auto& kv : (which == 0 ? S->ins : S->del))
		for (auto& sv : kv.second) { o.put<int64_t>(kv.first); o.put<int32_t>(sv.second); o.put<int32_t>((int32_t)sv.first.size()); o.b.insert(o.b.end(), sv.first.begin(), sv.first.end()); }
	return release(o, nbytes);
}
uint8_t* mco_breakpoints(void* h, int64_t* nbytes)
{
	State* S = (State*)h; Blob o;
	for (auto& kv : S->bp) { o.put<int64_t>(kv.first); o.put<int64_t>(kv.second); }
	return release(o, nbytes);
}
uint8_t* mco_sites(void* h, int which, int64_t* nbytes)
{
	State* S = (State*)h; Blob o;
	std::vector<std::pair<i64, i64> > v = which == 0 ? S->inv : S->tnl;
	std::stable_sort(v.begin(), v.end(), [](const std::pair<i64, i64>& x, const std::pair<i64, i64>& y) { return x.first < y.first; });
	for (auto& p : v) { o.put<int64_t>(p.first); o.put<int64_t>(p.second); }
	return release(o, nbytes);
}
void mco_bwt_search(void* h, const uint8_t* codes, int start, int stop, int* len, int* freq, uint64_t* loc)
{
	State* S = (State*)h; std::vector<uint8_t> c(codes, codes + stop);
	SearchResult r = bwt_search(S->ix, c, start, stop);
	*len = r.len; *freq = r.freq; for (int i = 0; i < r.freq; i++) loc[i] = r.loc[i];
}
int mco_align(int use_nw, int m, const char* s1, int n, const char* s2, char* o1, char* o2)
{
	std::string a(s1, m), b(s2, n);
	if (use_nw) nw_align(a, b); else ksw2_align(a, b);
	memcpy(o1, a.data(), a.size()); o1[a.size()] = 0; memcpy(o2, b.data(), b.size()); o2[b.size()] = 0;
	return a.size() == b.size() ? (int)a.size() : -1;
}
void mco_free(void* p) { free(p); }

} // extern "C"
