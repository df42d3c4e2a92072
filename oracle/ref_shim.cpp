// ORACLE / TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library this file is linked into.
//
// C wrappers around the UNMODIFIED reference objects (compiled from /root/reference/src by
// oracle/Makefile) so that Python tests can call the reference's own functions one by one and
// read back the global state the reference leaves behind.
//
// Two ways to run reads through the reference:
//   mcref_run_mapping()  calls the reference's own Mapping() (src/ReadMapping.cpp:689) on FASTQ
//                        files, i.e. the real thread body ReadMapping() with every quirk; only
//                        the globals (profile, maps, site vectors, counters) and the SAM file
//                        are observable afterwards.
//   mcref_map()          a driver loop written here that calls the reference's operator-level
//                        functions in the order of the thread body (src/ReadMapping.cpp:445-573)
//                        and serialises every read's AlnSummary / AlnCanVec / FragPairVec.  Its
//                        global side effects are checked against mcref_run_mapping() in
//                        tests/test_ref_shim.py, which is what licenses its per-read dumps.
#include "structure.h"

// non-static functions of the reference that structure.h does not declare
extern vector<FragPair_t> IdentifySimplePairs(int rlen, uint8_t* EncodeSeq);   // src/ReadMapping.cpp:125
extern void RemoveRedundantAlnCan(vector<AlnCan_t>& AlnCanVec);                // :228
extern int CheckPairedAlignmentDistance(int64_t EstiDistance, vector<AlnCan_t>& v1, vector<AlnCan_t>& v2); // :244
extern void MaskUnPairedAlnCan(vector<AlnCan_t>& v1, vector<AlnCan_t>& v2);    // :305
extern CoordinatePair_t GenCoordinatePair(vector<AlnCan_t>& v1, vector<AlnCan_t>& v2); // :361
extern void ResetPairedIdx(vector<AlnCan_t>& AlnCanVec);                       // :69
extern void EnCodeReadSeq(int rlen, char* seq, uint8_t* EncodeSeq);            // :404
extern bool CompByDiscordPos(const DiscordPair_t& p1, const DiscordPair_t& p2);
extern map<int64_t, uint16_t> BreakPointMap;                                   // src/AlignmentProfile.cpp:6
extern float MaxMisMatchRate, FrequencyThr;                                    // src/main.cpp:23
extern uint32_t avgReadLength;
extern int64_t iTotalReadNum, iTotalMappingNum, iTotalPairedNum, TotalPairedDistance, ReadLengthSum;
extern int64_t iAlignedBase, iTotalCoverage;
extern bwtint_t bwt_sa(bwtint_t k);
// src/VariantCalling.cpp: globals and thread bodies of the per-column scan
extern int* BlockDepthArr; extern int BlockNum; extern vector<Variant_t> VariantVec;
extern void* CalBlockReadDepth(void* arg); extern void* IdentifyVariants(void* arg); extern void RemoveConsecutiveGenomicVariant();

namespace {

struct Blob {
	vector<uint8_t> b;
	template <class T> void put(T v) { const uint8_t* p = (const uint8_t*)&v; b.insert(b.end(), p, p + sizeof(T)); }
	void bytes(const void* p, size_t n) { b.insert(b.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
};

uint8_t* blob_release(Blob& bl, int64_t* n)
{
	uint8_t* out = (uint8_t*)malloc(bl.b.size() ? bl.b.size() : 1);
	memcpy(out, bl.b.data(), bl.b.size());
	*n = (int64_t)bl.b.size();
	return out;
}

// state of the driver loop that the thread body keeps in locals (src/ReadMapping.cpp:418-427)
DiscordPair_t g_discord;          // deliberately persistent: reproduces the stale-gPos push of :502
bool g_discord_init = false;
vector<DiscordPair_t> g_inv, g_tnl;

void serialise_read(Blob& out, ReadItem_t& r)
{
	out.put<int32_t>(r.rlen);
	out.put<int32_t>(r.AlnSummary.score);
	out.put<int32_t>(r.AlnSummary.sub_score);
	out.put<int32_t>(r.AlnSummary.BestAlnCanIdx);
	out.put<int32_t>((int32_t)r.AlnCanVec.size());
	for (size_t c = 0; c < r.AlnCanVec.size(); c++)
	{
		AlnCan_t& a = r.AlnCanVec[c];
		out.put<int32_t>(a.score);
		out.put<int32_t>(a.score > 0 ? (a.orientation ? 1 : 0) : -1);
		out.put<int32_t>(a.PairedAlnCanIdx);
		// fragments of dead candidates are never consumed downstream (SamReport / UpdateProfile skip score==0)
		int32_t nf = a.score > 0 ? (int32_t)a.FragPairVec.size() : 0;
		out.put<int32_t>(nf);
		for (int32_t f = 0; f < nf; f++)
		{
			FragPair_t& fp = a.FragPairVec[f];
			out.put<int32_t>(fp.bSimple ? 1 : 0);
			out.put<int32_t>(fp.rPos);
			out.put<int64_t>(fp.gPos);
			out.put<int32_t>(fp.rLen);
			out.put<int32_t>(fp.gLen);
			int32_t al = fp.bSimple ? 0 : (int32_t)fp.aln1.length();
			out.put<int32_t>(al);
			if (al) { out.bytes(fp.aln1.data(), al); out.bytes(fp.aln2.data(), al); }
		}
	}
}

} // namespace

extern "C" {

// defaults of src/main.cpp:159-191 (the shim never runs main())
int mcref_load(const char* prefix)
{
	bGVCF = false; iPloidy = 2; iThreadNum = 1; bPairEnd = false; bDebugMode = false; bUnique = true; bFilter = false;
	NW_ALG = true; FastQFormat = true; bSAMoutput = false; bSAMFormat = true; bSomatic = false; bVCFoutput = true;
	gzCompressed = false; bMonomorphic = false;
	MaxClipSize = 5; MinCNVsize = 50; MaxPosDiff = 30; MinReadDepth = 20; iMaxDuplicate = 5; FragmentSize = 500;
	MinAlleleDepth = 5; FrequencyThr = 0.2; MinVarConfScore = 10; MinUnmappedSize = 50; MaxMisMatchRate = 0.05;
	sample_id = (char*)"unknown"; LogFileName = (char*)"/dev/null"; VcfFileName = (char*)"/dev/null";
	ObservGenomicPos = ObserveBegPos = ObserveEndPos = -1;
	RefSequence = RefFileName = SamFileName = KnownSiteFileName = NULL;
	IndexFileName = strdup(prefix);
	if (!CheckBWAIndexFiles(prefix)) return -1;
	RefIdx = bwa_idx_load(prefix);
	if (RefIdx == 0) return -2;
	Refbwt = RefIdx->bwt;
	iThreadNum = 4; RestoreReferenceInfo(); iThreadNum = 1;
	MappingRecordArr = new MappingRecord_t[GenomeSize]();
	pthread_mutex_init(&VarLock, NULL); pthread_mutex_init(&OutputLock, NULL);
	pthread_mutex_init(&LibraryLock, NULL); pthread_mutex_init(&ProfileLock, NULL);
	StartProcessTime = time(NULL);
	CmdLine = "mcref_shim";
	return 0;
}

int mcref_build_index(const char* fa, const char* prefix)
{
	extern int bwa_idx_build(const char* fa, const char* prefix);
	return bwa_idx_build(fa, prefix);
}

void mcref_set_params(int max_pos_diff, int max_clip, int max_dup, float maxmm, int nw_alg, int unique, int threads)
{
	MaxPosDiff = max_pos_diff; MaxClipSize = max_clip; iMaxDuplicate = (uint8_t)max_dup; MaxMisMatchRate = maxmm;
	NW_ALG = nw_alg != 0; bUnique = unique != 0; iThreadNum = threads;
}

void mcref_set_vc_flags(int monomorphic, int gvcf, int somatic, int filter)
{
	bMonomorphic = monomorphic != 0; bGVCF = gvcf != 0; bSomatic = somatic != 0; bFilter = filter != 0;
}

int64_t mcref_genome_size() { return GenomeSize; }
const char* mcref_refseq() { return RefSequence; }

// index internals, so tests can upload exactly what the reference holds in memory
const void* mcref_bwt_words(int64_t* n_words) { *n_words = (int64_t)Refbwt->bwt_size; return Refbwt->bwt; }
const void* mcref_sa(int64_t* n_sa) { *n_sa = (int64_t)Refbwt->n_sa; return Refbwt->sa; }
void mcref_bwt_meta(uint64_t out[8])
{
	out[0] = Refbwt->primary; for (int i = 0; i < 5; i++) out[1 + i] = Refbwt->L2[i];
	out[6] = Refbwt->seq_len; out[7] = (uint64_t)Refbwt->sa_intv;
}
int mcref_n_chrom() { return iChromsomeNum; }
void mcref_chrom(int i, int64_t* fwd_loc, int* len, char* name, int cap)
{
	*fwd_loc = ChromosomeVec[i].FowardLocation; *len = ChromosomeVec[i].len;
	strncpy(name, ChromosomeVec[i].name, cap - 1); name[cap - 1] = 0;
}

// ---- operator-level wrappers ------------------------------------------------------------------

// BWT_Search (src/bwt_search.cpp:121); loc must hold 50 entries
void mcref_bwt_search(const uint8_t* codes, int start, int stop, int* len, int* freq, uint64_t* loc)
{
	bwtSearchResult_t r = BWT_Search((uint8_t*)codes, start, stop);
	*len = r.len; *freq = r.freq;
	for (int i = 0; i < r.freq; i++) loc[i] = r.LocArr[i];
	if (r.LocArr) delete[] r.LocArr;
}

uint64_t mcref_bwt_sa(uint64_t k) { return bwt_sa(k); }

// nw_alignment (src/nw_alignment.cpp:18) / ksw2_alignment (src/ksw2_alignment.cpp:250).
// Returns the aligned length; o1/o2 must hold m+n+1 bytes.
int mcref_align(int use_nw, int m, const char* s1, int n, const char* s2, char* o1, char* o2)
{
	string a(s1, m), b(s2, n);
	if (use_nw) nw_alignment(m, a, n, b); else ksw2_alignment(m, a, n, b);
	memcpy(o1, a.data(), a.size()); memcpy(o2, b.data(), b.size());
	o1[a.size()] = 0; o2[b.size()] = 0;
	return (int)a.size() == (int)b.size() ? (int)a.size() : -(int)a.size();
}

// seeds + clusters of one read: IdentifySimplePairs + SimplePairClustering (src/ReadMapping.cpp:125,194)
uint8_t* mcref_seed_cluster(const char* seq, int rlen, int64_t* nbytes)
{
	Blob out;
	vector<uint8_t> codes(rlen ? rlen : 1);
	EnCodeReadSeq(rlen, (char*)seq, codes.data());
	vector<FragPair_t> sp = IdentifySimplePairs(rlen, codes.data());
	out.put<int32_t>((int32_t)sp.size() - 1); // without the sentinel
	for (size_t i = 0; i + 1 < sp.size(); i++) { out.put<int32_t>(sp[i].rPos); out.put<int64_t>(sp[i].gPos); out.put<int32_t>(sp[i].rLen); }
	vector<AlnCan_t> cv = SimplePairClustering(rlen, sp);
	out.put<int32_t>((int32_t)cv.size());
	for (size_t c = 0; c < cv.size(); c++)
	{
		out.put<int32_t>(cv[c].score); out.put<int32_t>((int32_t)cv[c].FragPairVec.size());
		for (size_t f = 0; f < cv[c].FragPairVec.size(); f++)
		{
			out.put<int32_t>(cv[c].FragPairVec[f].rPos); out.put<int64_t>(cv[c].FragPairVec[f].gPos); out.put<int32_t>(cv[c].FragPairVec[f].rLen);
		}
	}
	return blob_release(out, nbytes);
}

// ---- whole-pipeline driver --------------------------------------------------------------------

void mcref_reset_state()
{
	if (MappingRecordArr) memset((void*)MappingRecordArr, 0, sizeof(MappingRecord_t) * GenomeSize);
	InsertSeqMap.clear(); DeleteSeqMap.clear(); BreakPointMap.clear();
	InversionSiteVec.clear(); TranslocationSiteVec.clear(); g_inv.clear(); g_tnl.clear();
	iTotalReadNum = iTotalMappingNum = iTotalPairedNum = TotalPairedDistance = ReadLengthSum = 0;
	iAlignedBase = iTotalCoverage = 0;
	avgDist = 1000; avgReadLength = 0; g_discord_init = false;
}

// operator-level tests (AlignmentRescue with a given EstiDistance): the value the next chunk computes EstiDistance from
void mcref_set_avg_dist(uint32_t v) { avgDist = v; }

// Processes `n_reads` reads (mates adjacent when paired) in chunks of ReadChunkSize exactly as one
// reference thread would (src/ReadMapping.cpp:432-624), mutating the reference's global state.
// seq: concatenated ASCII reads, off[n_reads+1].  Returns a malloc'd blob: for every read the record
// written by serialise_read(); for paired input additionally, after all reads, one int32 EstiDistance
// per chunk (the (int)(avgDist*1.5) the chunk was processed with).
uint8_t* mcref_map(int64_t n_reads, const char* seq, const int64_t* off, int paired, int update_profile, int64_t* nbytes)
{
	Blob out;
	vector<int32_t> est_per_chunk;
	AlnSummary_t blank; blank.score = blank.sub_score = 0; blank.BestAlnCanIdx = -1;
	bPairEnd = paired != 0;
	if (!g_discord_init) { g_discord.gPos = 0; g_discord.dist = 0; g_discord_init = true; } // :420 leaves it uninitialised; the tests avoid reading the stale value before a first in-range pair

	for (int64_t base = 0; base < n_reads; base += ReadChunkSize)
	{
		int ReadNum = (int)min<int64_t>(ReadChunkSize, n_reads - base);
		vector<ReadItem_t> R(ReadNum);
		for (int i = 0; i < ReadNum; i++)
		{
			int64_t b = off[base + i], e = off[base + i + 1];
			R[i].rlen = (int)(e - b);
			R[i].seq = new char[R[i].rlen + 1]; memcpy(R[i].seq, seq + b, R[i].rlen); R[i].seq[R[i].rlen] = 0;
			R[i].qual = NULL; R[i].header = NULL;
		}
		int MappedNum = 0, PairedNum = 0; int64_t myDist = 0, myLen = 0;
		if (paired && ReadNum % 2 == 0)
		{
			int est = (int)(avgDist * 1.5);
			est_per_chunk.push_back(est);
			for (int i = 0, j = 1; i != ReadNum; i += 2, j += 2)
			{
				vector<uint8_t> codes(R[i].rlen ? R[i].rlen : 1);
				EnCodeReadSeq(R[i].rlen, R[i].seq, codes.data());
				vector<FragPair_t> sp = IdentifySimplePairs(R[i].rlen, codes.data());
				R[i].AlnCanVec = SimplePairClustering(R[i].rlen, sp);
				ReverseOrientation(&R[j]);
				codes.assign(R[j].rlen ? R[j].rlen : 1, 0);
				EnCodeReadSeq(R[j].rlen, R[j].seq, codes.data());
				sp = IdentifySimplePairs(R[j].rlen, codes.data());
				R[j].AlnCanVec = SimplePairClustering(R[j].rlen, sp);
				R[i].AlnSummary = blank; R[j].AlnSummary = blank;
				ResetPairedIdx(R[i].AlnCanVec); ResetPairedIdx(R[j].AlnCanVec);
				int n = CheckPairedAlignmentDistance(est, R[i].AlnCanVec, R[j].AlnCanVec);
				if (n == 0) n = AlignmentRescue(est, R[i], R[j]);
				if (n == 0) { RemoveRedundantAlnCan(R[i].AlnCanVec); RemoveRedundantAlnCan(R[j].AlnCanVec); }
				else MaskUnPairedAlnCan(R[i].AlnCanVec, R[j].AlnCanVec);
				if (ProduceReadAlignment(R[i])) MappedNum++;
				if (ProduceReadAlignment(R[j])) MappedNum++;
				CoordinatePair_t cp = GenCoordinatePair(R[i].AlnCanVec, R[j].AlnCanVec);
				if (cp.dist != 0 && cp.gPos1 != -1 && cp.gPos2 != -1)
				{
					// pair classification, src/ReadMapping.cpp:486-532 (including the brace-less push of :502)
					if (cp.gPos1 < GenomeSize && cp.gPos2 >= GenomeSize)
					{
						if (bVCFoutput)
						{
							g_discord.dist = abs(TwoGenomeSize - cp.gPos1 - cp.gPos2);
							if (g_discord.dist > 1000 && g_discord.dist < 10000000) { g_discord.gPos = cp.gPos1; g_inv.push_back(g_discord); }
						}
					}
					else if (cp.gPos1 >= GenomeSize && cp.gPos2 < GenomeSize)
					{
						if (bVCFoutput)
						{
							g_discord.dist = abs(TwoGenomeSize - cp.gPos1 - cp.gPos2);
							if (g_discord.dist > 1000 && g_discord.dist < 10000000) g_discord.gPos = cp.gPos2;
							g_inv.push_back(g_discord);
						}
					}
					else if (cp.dist > 1000)
					{
						if (bVCFoutput)
						{
							g_discord.dist = cp.dist;
							if (cp.gPos1 < GenomeSize && cp.gPos2 < GenomeSize)
							{
								g_discord.gPos = cp.gPos1; g_tnl.push_back(g_discord);
								g_discord.gPos = cp.gPos2; g_tnl.push_back(g_discord);
							}
							else if (cp.gPos1 >= GenomeSize && cp.gPos2 >= GenomeSize)
							{
								g_discord.gPos = TwoGenomeSize - cp.gPos1; g_tnl.push_back(g_discord);
								g_discord.gPos = TwoGenomeSize - cp.gPos2; g_tnl.push_back(g_discord);
							}
						}
					}
					else
					{
						myLen += R[i].rlen; myLen += R[j].rlen; PairedNum++; myDist += cp.dist;
					}
				}
			}
			iTotalReadNum += ReadNum; iTotalMappingNum += MappedNum; iTotalPairedNum += PairedNum;
			TotalPairedDistance += myDist; ReadLengthSum += myLen;
			if (iTotalPairedNum > 1000) avgDist = (int)(1. * TotalPairedDistance / iTotalPairedNum + .5);
			if (update_profile)
			{
				for (int i = 0; i != ReadNum; i++)
				{
					if (R[i].AlnSummary.score == 0) continue;
					if (CheckAlnNumber(R[i].AlnCanVec) == 1) UpdateProfile((i % 2 == 0), &R[i], R[i].AlnCanVec);
					else UpdateMultiHitCount(&R[i], R[i].AlnCanVec);
				}
			}
		}
		else
		{
			for (int i = 0; i != ReadNum; i++)
			{
				vector<uint8_t> codes(R[i].rlen ? R[i].rlen : 1);
				EnCodeReadSeq(R[i].rlen, R[i].seq, codes.data());
				vector<FragPair_t> sp = IdentifySimplePairs(R[i].rlen, codes.data());
				R[i].AlnSummary = blank; R[i].AlnCanVec = SimplePairClustering(R[i].rlen, sp);
				RemoveRedundantAlnCan(R[i].AlnCanVec);
				if (ProduceReadAlignment(R[i])) MappedNum++;
			}
			iTotalReadNum += ReadNum; iTotalMappingNum += MappedNum;
			if (update_profile)
			{
				for (int i = 0; i != ReadNum; i++)
				{
					if (R[i].AlnSummary.score == 0) continue;
					if (CheckAlnNumber(R[i].AlnCanVec) == 1) UpdateProfile(true, &R[i], R[i].AlnCanVec);
					else UpdateMultiHitCount(&R[i], R[i].AlnCanVec);
				}
			}
		}
		for (int i = 0; i < ReadNum; i++) { serialise_read(out, R[i]); delete[] R[i].seq; }
	}
	for (size_t i = 0; i < est_per_chunk.size(); i++) out.put<int32_t>(est_per_chunk[i]);
	return blob_release(out, nbytes);
}

// thread-end merge of the site vectors (src/ReadMapping.cpp:627-644) for the driver loop above
void mcref_finish_sites()
{
	sort(g_tnl.begin(), g_tnl.end(), CompByDiscordPos); sort(g_inv.begin(), g_inv.end(), CompByDiscordPos);
	TranslocationSiteVec = g_tnl; InversionSiteVec = g_inv;
}

// the real thing: Mapping() on FASTQ/FASTA files with `threads` reference threads
int mcref_run_mapping(const char* f1, const char* f2, const char* sam_path, int threads, int vcf)
{
	ReadFileNameVec1.clear(); ReadFileNameVec2.clear();
	ReadFileNameVec1.push_back(f1); if (f2 && f2[0]) ReadFileNameVec2.push_back(f2);
	bPairEnd = (f2 && f2[0]);
	bSAMoutput = sam_path && sam_path[0]; bSAMFormat = true; SamFileName = bSAMoutput ? strdup(sam_path) : NULL;
	bVCFoutput = vcf != 0; iThreadNum = threads;
	StartProcessTime = time(NULL);
	Mapping();
	return 0;
}

int mcref_variant_calling(const char* vcf_path)
{
	VcfFileName = strdup(vcf_path);
	VariantCalling();
	return 0;
}

// The per-column part of VariantCalling() (src/VariantCalling.cpp:706-717): CalBlockReadDepth, IdentifyVariants on one
// thread (the reference forces iThreadNum = 1 there) and RemoveConsecutiveGenomicVariant, on the profile the mapping
// left behind.  Records: int64 gPos, int32 VarType, DP, AD_ref, AD_alt, GenoType, qscore, alt_len, bytes; then int32
// BlockNum and BlockDepthArr.  Fields the reference leaves stale for a type are serialised as they are.
uint8_t* mcref_variant_scan(int min_ad, float freq_thr, int somatic, int gvcf, int mono, int ploidy, int min_cnv, int min_gap, int64_t* nbytes)
{
	MinAlleleDepth = min_ad; FrequencyThr = freq_thr; bSomatic = somatic != 0; bGVCF = gvcf != 0; bMonomorphic = mono != 0;
	iPloidy = ploidy; MinCNVsize = min_cnv; MinUnmappedSize = min_gap;
	if (bGVCF && bMonomorphic) bGVCF = false;   // src/main.cpp:322
	const int saved = iThreadNum; iThreadNum = 1; int tid = 0;
	BlockNum = (int)(GenomeSize / 100); if (((int64_t)BlockNum * 100) < GenomeSize) BlockNum += 1;
	BlockDepthArr = new int[BlockNum]();
	CalBlockReadDepth(&tid);
	VariantVec.clear();
	IdentifyVariants(&tid);
	if (bGVCF && VariantVec.size() > 0) RemoveConsecutiveGenomicVariant();
	Blob out;
	out.put<int64_t>((int64_t)VariantVec.size());
	for (size_t i = 0; i < VariantVec.size(); i++)
	{
		Variant_t& v = VariantVec[i];
		out.put<int64_t>(v.gPos); out.put<int32_t>(v.VarType); out.put<int32_t>(v.DP); out.put<int32_t>(v.AD_ref); out.put<int32_t>(v.AD_alt);
		out.put<int32_t>(v.GenoType); out.put<int32_t>(v.qscore); out.put<int32_t>((int32_t)v.ALTstr.size());
		out.bytes(v.ALTstr.data(), v.ALTstr.size());
	}
	out.put<int32_t>(BlockNum);
	out.bytes(BlockDepthArr, (size_t)BlockNum * 4);
	delete[] BlockDepthArr; BlockDepthArr = NULL; VariantVec.clear(); iThreadNum = saved;
	return blob_release(out, nbytes);
}

void mcref_counters(int64_t out[8])
{
	out[0] = iTotalReadNum; out[1] = iTotalMappingNum; out[2] = iTotalPairedNum; out[3] = TotalPairedDistance;
	out[4] = ReadLengthSum; out[5] = avgDist; out[6] = avgReadLength; out[7] = FragmentSize;
}

// profile as 10 int32 per position: A C G T multi_hit readCount F1 R2 F2 R1
void mcref_profile(int64_t beg, int64_t end, int32_t* out)
{
	for (int64_t g = beg; g < end; g++)
	{
		MappingRecord_t& m = MappingRecordArr[g]; int32_t* o = out + (g - beg) * 10;
		o[0] = m.A; o[1] = m.C; o[2] = m.G; o[3] = m.T; o[4] = m.multi_hit; o[5] = m.readCount;
		o[6] = m.F1; o[7] = m.R2; o[8] = m.F2; o[9] = m.R1;
	}
}
const void* mcref_profile_raw() { return MappingRecordArr; }

// which: 0 InsertSeqMap, 1 DeleteSeqMap.  Records: int64 pos, int32 count, int32 len, bytes
uint8_t* mcref_indels(int which, int64_t* nbytes)
{
	Blob out;
	map<int64_t, map<string, uint16_t> >& M = which == 0 ? InsertSeqMap : DeleteSeqMap;
	for (map<int64_t, map<string, uint16_t> >::iterator it = M.begin(); it != M.end(); it++)
		for (map<string, uint16_t>::iterator s = it->second.begin(); s != it->second.end(); s++)
		{
			out.put<int64_t>(it->first); out.put<int32_t>(s->second); out.put<int32_t>((int32_t)s->first.size());
			out.bytes(s->first.data(), s->first.size());
		}
	return blob_release(out, nbytes);
}

// int64 pos, int64 count pairs
uint8_t* mcref_breakpoints(int64_t* nbytes)
{
	Blob out;
	for (map<int64_t, uint16_t>::iterator it = BreakPointMap.begin(); it != BreakPointMap.end(); it++) { out.put<int64_t>(it->first); out.put<int64_t>(it->second); }
	return blob_release(out, nbytes);
}

// which: 0 InversionSiteVec, 1 TranslocationSiteVec.  int64 gPos, int64 dist pairs in vector order
uint8_t* mcref_sites(int which, int64_t* nbytes)
{
	Blob out;
	vector<DiscordPair_t>& V = which == 0 ? InversionSiteVec : TranslocationSiteVec;
	for (size_t i = 0; i < V.size(); i++) { out.put<int64_t>(V[i].gPos); out.put<int64_t>(V[i].dist); }
	return blob_release(out, nbytes);
}

void mcref_free(void* p) { free(p); }

} // extern "C"
