// Drop-in replacement for the reference's hot-path translation units
//   ReadMapping.cpp bwt_search.cpp ReadAlignment.cpp nw_alignment.cpp ksw2_alignment.cpp AlignmentProfile.cpp
//   AlignmentRescue.cpp KmerAnalysis.cpp
// It is compiled against the reference's own (unchanged) src/structure.h and linked with the reference's unchanged
// main.o GetData.o VariantCalling.o SamReport.o tools.o bwt_index.o (+ BWT_Index), giving a MapCaller binary with the same
// command line whose mapping runs on the GPU through libmapcaller_b200.so.  It provides exactly the
// symbols the kept objects import from the replaced ones: Mapping(), CompByDiscordPos(),
// avgDist, avgReadLength, BreakPointMap, InsertSeqMap, DeleteSeqMap, InversionSiteVec, TranslocationSiteVec.
//
// Plain FASTQ files never pass through the reference's reader: raw file blocks go to the device, which finds the records,
// maps them and - with -sam - assembles the SAM text (mc_ingest_fastq / mc_map_staged / mc_sam_text).  .gz and FASTA input
// still comes through the reference's GetNextChunk / gzGetNextChunk; its SAM lines are then made by the reference's
// SamReport.o from the downloaded candidates, or - MC_B200_DEVICE_SAM=1 - printed from the device's record fields
// (mc_sam_records).  Everything after Mapping() is the reference's, except IdentifyVariants (mc_variant_scan, see below).
#include "structure.h"
#include "mapcaller_b200.h"

// ---- symbols the kept objects import ----------------------------------------------------------------------
vector<DiscordPair_t> InversionSiteVec, TranslocationSiteVec;
map<int64_t, map<string, uint16_t> > InsertSeqMap, DeleteSeqMap;
map<int64_t, uint16_t> BreakPointMap;
uint32_t avgCov, avgReadLength, avgDist = 1000;
int64_t iTotalReadNum = 0, iTotalMappingNum = 0, iTotalPairedNum = 0, iAlignedBase = 0, iTotalCoverage = 0, TotalPairedDistance = 0, ReadLengthSum = 0;
extern float MaxMisMatchRate, FrequencyThr;

bool CompByDiscordPos(const DiscordPair_t& p1, const DiscordPair_t& p2) { return p1.gPos < p2.gPos; }

// IdentifyVariants (src/VariantCalling.cpp:550-680) - the reference's forced-single-thread scan over every column of
// MappingRecordArr - is answered by mc_variant_scan on the device-resident profile: Mapping() runs the scan while the context
// still exists and parks the records here; the Makefile weakens the symbol in the kept VariantCalling.o (objcopy
// --weaken-symbol), so the unchanged VariantCalling() calls THIS definition and finds VariantVec filled.  Everything after
// it (RemoveConsecutiveGenomicVariant, structural variants, filters, the VCF text) is the reference's own code.
extern vector<Variant_t> VariantVec;
static vector<Variant_t> g_scanned_variants;
void* IdentifyVariants(void*) { VariantVec.swap(g_scanned_variants); return (void*)(1); }

namespace {

const int kBatchChunks = 2500;   // 200-read chunks per GPU batch (500 k reads)

void die(const char* what) { fprintf(stderr, "\n[mapcaller_b200] %s: %s\n", what, mc_last_error()); exit(1); }

struct Library { FILE *f1, *f2; gzFile g1, g2; bool sep, gz; };

// one GPU batch: up to kBatchChunks chunks pulled with the reference's own reader
int64_t pull_batch(Library& lib, vector<ReadItem_t>& reads)
{
	reads.clear();
	vector<ReadItem_t> chunk(ReadChunkSize);
	for (int c = 0; c < kBatchChunks; c++)
	{
		int n = lib.gz ? gzGetNextChunk(lib.sep, lib.g1, lib.g2, chunk.data()) : GetNextChunk(lib.sep, lib.f1, lib.f2, chunk.data());
		if (n == 0) break;
		reads.insert(reads.end(), chunk.begin(), chunk.begin() + n);
		if (n < ReadChunkSize) break;
	}
	return (int64_t)reads.size();
}

void free_reads(vector<ReadItem_t>& reads)
{
	for (size_t i = 0; i < reads.size(); i++) { delete[] reads[i].header; delete[] reads[i].seq; delete[] reads[i].qual; }
	reads.clear();
}

// AlnSummary / AlnCanVec of one read from the flat batch result (inverse of the library's CSR layout)
void rebuild(ReadItem_t& rd, const mc_batch_out& out, int64_t r)
{
	const mc_read_out& ro = out.reads[r];
	rd.AlnSummary.score = ro.score; rd.AlnSummary.sub_score = ro.sub_score; rd.AlnSummary.BestAlnCanIdx = ro.best_idx;
	rd.AlnCanVec.clear(); rd.AlnCanVec.resize(ro.n_cand);
	for (int k = 0; k < ro.n_cand; k++)
	{
		const mc_cand_out& co = out.cands[ro.cand_begin + k];
		AlnCan_t& ac = rd.AlnCanVec[k];
		ac.score = co.score; ac.SamFlag = 0; ac.orientation = co.orientation == 1; ac.PairedAlnCanIdx = co.paired_idx;
		ac.FragPairVec.resize(co.n_frag);
		for (int f = 0; f < co.n_frag; f++)
		{
			const mc_frag_out& fo = out.frags[co.frag_begin + f];
			FragPair_t& fp = ac.FragPairVec[f];
			fp.bSimple = fo.bSimple != 0; fp.rPos = fo.rPos; fp.gPos = fo.gPos; fp.rLen = fo.rLen; fp.gLen = fo.gLen; fp.PosDiff = fo.gPos - fo.rPos;
			if (!fp.bSimple && fo.aln_len > 0)
			{
				fp.aln1.assign((const char*)out.aln + fo.aln_off, fo.aln_len);
				fp.aln2.assign((const char*)out.aln + fo.aln_off + fo.aln_cap, fo.aln_len);
			}
		}
	}
}

// One SAM line from the record the device computed (mc_sam_records): the text of GeneratePairedSamStream /
// GenerateSingleSamStream (src/SamReport.cpp:332,351,385-434) with QNAME / SEQ / QUAL from the read as the reader delivered it.
void print_sam_line(FILE* out, const ReadItem_t& rd, const mc_sam_rec& r, const uint8_t* cigar_arena, string& seq, string& qual)
{
	if (r.flag < 0) return;
	seq.assign(rd.seq, rd.rlen);
	if (FastQFormat) qual.assign(rd.qual, rd.rlen); else qual = "*";
	if (r.reverse)
	{
		seq.resize(rd.rlen + 1); GetComplementarySeq(rd.rlen, rd.seq, &seq[0]); seq.resize(rd.rlen);
		if (FastQFormat) reverse(qual.begin(), qual.end());
	}
	if (r.chrom < 0) { fprintf(out, "%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\tAS:i:0\tXS:i:0\n", rd.header, r.flag, seq.c_str(), qual.c_str()); return; }
	fprintf(out, "%s\t%d\t%s\t%lld\t%d\t%.*s\t", rd.header, r.flag, ChromosomeVec[r.chrom].name, (long long)r.pos, r.mapq, r.cigar_len, (const char*)cigar_arena + r.cigar_off);
	if (r.has_mate) fprintf(out, "=\t%lld\t%d", (long long)r.mate_pos, r.tlen); else fprintf(out, "*\t0\t0");
	fprintf(out, "\t%s\t%s\tNM:i:%d\tAS:i:%d\tXS:i:%d\n", seq.c_str(), qual.c_str(), r.nm, r.as, r.xs);
}

} // namespace

void Mapping()
{
	FILE* sam_out = NULL;
	if (bSAMoutput && !bSAMFormat) { fprintf(stderr, "Error! -bam is not supported by the GPU build; use -sam\n"); exit(1); }
	if (bSAMoutput && SamFileName != NULL) sam_out = strcmp(SamFileName, "-") == 0 ? fopen("/dev/stdout", "w") : fopen(SamFileName, "w");
	if (bSAMoutput)
	{
		fprintf(sam_out, "@PG\tID:MapCaller\tPN:MapCaller\tVN:%s\n", VersionStr);
		for (int i = 0; i < iChromsomeNum; i++) fprintf(sam_out, "@SQ\tSN:%s\tLN:%d\n", ChromosomeVec[i].name, ChromosomeVec[i].len);
	}

	// the index the reference has just loaded, handed over without copying
	vector<int32_t> clen(iChromsomeNum); vector<const char*> cname(iChromsomeNum);
	for (int i = 0; i < iChromsomeNum; i++) { clen[i] = ChromosomeVec[i].len; cname[i] = ChromosomeVec[i].name; }
	mc_index_view v; memset(&v, 0, sizeof(v));
	v.bwt = Refbwt->bwt; v.bwt_size = Refbwt->bwt_size; v.primary = Refbwt->primary; for (int i = 0; i < 5; i++) v.L2[i] = Refbwt->L2[i];
	v.seq_len = Refbwt->seq_len; v.sa = Refbwt->sa; v.n_sa = Refbwt->n_sa; v.sa_intv = Refbwt->sa_intv; v.pac = RefIdx->pac; v.genome_size = GenomeSize;
	v.n_chrom = iChromsomeNum; v.chrom_len = clen.data(); v.chrom_name = cname.data();
	mc_index* idx = NULL; if (mc_index_wrap(&v, &idx)) die("mc_index_wrap");

	// MC_B200_DEVICE_SAM=1: flags, positions, MAPQ, CIGAR and mate fields come from the device (mc_sam_records) and only the
	// text is printed here, instead of downloading every candidate / fragment / alignment string and running the reference's
	// SamReport.o over them.  One line per read, i.e. not with -m (bUnique = false).
	const bool device_sam = bSAMoutput && bUnique && getenv("MC_B200_DEVICE_SAM") != NULL;
	mc_ctx* ctx = NULL;
	for (int lib_id = 0; lib_id < (int)ReadFileNameVec1.size(); lib_id++)
	{
		Library lib; memset(&lib, 0, sizeof(lib));
		const string& n1 = ReadFileNameVec1[lib_id];
		lib.gz = gzCompressed = n1.substr(n1.find_last_of('.') + 1) == "gz";
		FastQFormat = CheckReadFormat(n1.c_str());
		if (lib.gz) lib.g1 = gzopen(n1.c_str(), "rb"); else lib.f1 = fopen(n1.c_str(), "r");
		if (ReadFileNameVec1.size() == ReadFileNameVec2.size())
		{
			lib.sep = bPairEnd = true;
			const string& n2 = ReadFileNameVec2[lib_id];
			if (FastQFormat != CheckReadFormat(n2.c_str())) { fprintf(stderr, "Error! %s and %s are with different format...\n", n1.c_str(), n2.c_str()); continue; }
			if (lib.gz) lib.g2 = gzopen(n2.c_str(), "rb"); else lib.f2 = fopen(n2.c_str(), "r");
		}
		if (lib.f1 == NULL && lib.g1 == NULL) continue;
		if (lib.sep && lib.f2 == NULL && lib.g2 == NULL) continue;

		if (ctx == NULL)   // bPairEnd is only known once the first library is open
		{
			mc_params p; mc_params_default(&p);
			p.paired = bPairEnd; p.alg_ksw2 = !NW_ALG; p.max_pos_diff = MaxPosDiff; p.max_clip = MaxClipSize; p.max_dup = iMaxDuplicate;
			p.max_mismatch_rate = MaxMisMatchRate; p.update_profile = bVCFoutput; p.want_alignments = bSAMoutput && !device_sam;
			if (mc_ctx_create(idx, &p, &ctx)) die("mc_ctx_create");
		}
		if (mc_begin_library(ctx)) die("mc_begin_library");   // every -f library restarts the 200-read chunk grid; profile, totals and avgDist carry over
		// No SAM wanted and plain FASTQ on disk: nothing of the reference's reader is needed - raw file blocks go to the GPU, which
		// finds the records itself (mc_ingest_fastq = GetNextEntry / GetNextChunk, src/GetData.cpp:32-99) and maps them.
		// With -sam the lines are assembled on the device as well (mc_sam_text: QNAME / SEQ / QUAL from the FASTQ text kept in the
		// slot, -m included) and written with one fwrite per batch; MC_B200_HOST_SAM=1 keeps the reference's reader + SamReport.o.
		// .gz FASTQ takes the same road: the two mate files are inflated block-wise by zlib (gzread, one host thread per file)
		// and the device parses the inflated text - no per-line gzgets, no per-read allocation (gzGetNextEntry,
		// src/GetData.cpp:101-131; its 1024-byte line buffer is the one difference: longer lines are not cut here).
		// FASTA reads take it as well when every record keeps its bases on one line (mc_fastq_in.format = 1); the device checks that,
		// and a file with wrapped records is handed back to the reference's reader before anything has been mapped from it.
		bool device_reader = !bSAMoutput || getenv("MC_B200_HOST_SAM") == NULL;
		if (device_reader)
		{
			if (lib.gz) { gzbuffer(lib.g1, 1 << 20); if (lib.g2) gzbuffer(lib.g2, 1 << 20); }
			auto read_block = [&lib](int which, uint8_t* dst, size_t n) -> size_t {
				if (!lib.gz) return fread(dst, 1, n, which ? lib.f2 : lib.f1);
				size_t got = 0;
				while (got < n) { const int g = gzread(which ? lib.g2 : lib.g1, dst + got, (unsigned)std::min<size_t>(n - got, 1u << 30)); if (g <= 0) break; got += (size_t)g; }
				return got;
			};
			// MC_B200_FASTQ_BLOCK (bytes) overrides the 64 MiB block size - the tests use it to cross many block boundaries with small files
			const size_t BLK = getenv("MC_B200_FASTQ_BLOCK") ? (size_t)atoll(getenv("MC_B200_FASTQ_BLOCK")) : (size_t)64 << 20;
			vector<uint8_t> b1, b2; size_t have1 = 0, have2 = 0; bool eof1 = false, eof2 = !lib.sep, force_final = false, first_block = true;
			for (;;)
			{
				// a file that reached its end just stops growing; its carried-over records are still consumed block by block
				size_t g2 = 0; pthread_t th2; bool threaded = false;
				struct Job { decltype(read_block)* rd; uint8_t* dst; size_t n, got; } job = {&read_block, NULL, BLK, 0};
				if (!eof2)
				{
					b2.resize(have2 + BLK); job.dst = b2.data() + have2;
					// the second mate file is read (inflated) by a thread of its own while this one reads the first
					threaded = !eof1 && pthread_create(&th2, NULL, [](void* p) -> void* { Job* j = (Job*)p; j->got = (*j->rd)(1, j->dst, j->n); return NULL; }, &job) == 0;
					if (!threaded) g2 = read_block(1, job.dst, BLK);
				}
				if (!eof1) { b1.resize(have1 + BLK); size_t g = read_block(0, b1.data() + have1, BLK); have1 += g; eof1 = g < BLK; }
				if (threaded) { pthread_join(th2, NULL); g2 = job.got; }
				if (!eof2) { have2 += g2; eof2 = g2 < BLK; }
				const bool last = (eof1 && eof2) || force_final;   // only then may a ragged tail (not a multiple of 200 reads) be mapped
				mc_fastq_in fi; memset(&fi, 0, sizeof(fi));
				fi.text1 = b1.data(); fi.len1 = (int64_t)have1; fi.text2 = lib.sep ? b2.data() : NULL; fi.len2 = (int64_t)have2; fi.final_block = last;
				fi.format = FastQFormat ? 0 : 1;
				mc_fastq_out fo;
				if (mc_ingest_fastq(ctx, &fi, 0, &fo))
				{
					if (FastQFormat || !first_block) die("mc_ingest_fastq");
					device_reader = false;          // wrapped FASTA records: back to the start, the reference's reader takes the library
					break;
				}
				first_block = false;
				if (fo.n_reads > 0)
				{
					mc_batch_out out;
					if (mc_map_staged(ctx, 0, &out)) die("mc_map_staged");
					if (bSAMoutput)
					{
						const uint8_t* text; int64_t n_bytes;
						if (mc_sam_text(ctx, 0, !bUnique, &text, &n_bytes)) die("mc_sam_text");
						if (n_bytes > 0 && fwrite(text, 1, (size_t)n_bytes, sam_out) != (size_t)n_bytes) { fprintf(stderr, "\nError! cannot write %s\n", SamFileName); exit(1); }
					}
					mc_totals t; mc_get_totals(ctx, &t);
					fprintf(stderr, "\r%lld %s reads have been processed in %lld seconds...", (long long)t.total_reads, (bPairEnd ? "paired-end" : "singled-end"), (long long)(time(NULL) - StartProcessTime));
				}
				if (last) break;
				memmove(b1.data(), b1.data() + fo.consumed1, have1 - (size_t)fo.consumed1); have1 -= (size_t)fo.consumed1;
				if (lib.sep) { memmove(b2.data(), b2.data() + fo.consumed2, have2 - (size_t)fo.consumed2); have2 -= (size_t)fo.consumed2; }
				if (lib.sep && eof1 != eof2)
				{
					// mate files of unequal size: the shorter one ends the library (the reference's reader stops there too). Once the
					// file still open has buffered more whole records than the ended one has left, nothing else can pair up.
					const int64_t used = fo.n_reads / 2, left1 = fo.records1 - used, left2 = fo.records2 - used;
					if (eof1 ? left2 > left1 : left1 > left2) force_final = true;
				}
			}
			if (device_reader)
			{
				if (lib.gz) { gzclose(lib.g1); if (lib.g2) gzclose(lib.g2); } else { fclose(lib.f1); if (lib.f2) fclose(lib.f2); }
				continue;
			}
			if (lib.gz) { gzrewind(lib.g1); if (lib.g2) gzrewind(lib.g2); } else { rewind(lib.f1); if (lib.f2) rewind(lib.f2); }
		}
		vector<ReadItem_t> reads; vector<uint8_t> seq; vector<int64_t> off; vector<string> sam;
		while (pull_batch(lib, reads) > 0)
		{
			int64_t n = (int64_t)reads.size();
			if (bPairEnd && (n & 1)) { free_reads(reads); break; }   // the reference silently maps an odd chunk as single-end; not supported here
			off.assign(1, 0); seq.clear();
			for (int64_t i = 0; i < n; i++) { seq.insert(seq.end(), reads[i].seq, reads[i].seq + reads[i].rlen); off.push_back((int64_t)seq.size()); }
			mc_batch_in in; in.n_reads = n; in.seq = seq.data(); in.seq_off = off.data();
			mc_batch_out out;
			if (mc_map_batch(ctx, &in, &out)) die("mc_map_batch");
			mc_totals t; mc_get_totals(ctx, &t);
			fprintf(stderr, "\r%lld %s reads have been processed in %lld seconds...", (long long)t.total_reads, (bPairEnd ? "paired-end" : "singled-end"), (long long)(time(NULL) - StartProcessTime));
			if (device_sam)
			{
				const mc_sam_rec* sr; int64_t ns; const uint8_t* cig; string sq, ql;
				if (mc_sam_records(ctx, &sr, &ns, &cig)) die("mc_sam_records");
				for (int64_t i = 0; i < n; i++) print_sam_line(sam_out, reads[i], sr[i], cig, sq, ql);
				fflush(sam_out);
			}
			else if (bSAMoutput)
			{
				sam.clear();
				for (int64_t i = 0; i < n; i++) rebuild(reads[i], out, i);
				if (bPairEnd) for (int64_t i = 0; i < n; i += 2) { ReverseOrientation(&reads[i + 1]); GeneratePairedSamStream(reads[i], reads[i + 1], sam); }
				else for (int64_t i = 0; i < n; i++) GenerateSingleSamStream(reads[i], sam);
				for (size_t i = 0; i < sam.size(); i++) fprintf(sam_out, "%s\n", sam[i].c_str());
				fflush(sam_out);
			}
			free_reads(reads);
		}
		if (lib.gz) { if (lib.g1) gzclose(lib.g1); if (lib.g2) gzclose(lib.g2); } else { if (lib.f1) fclose(lib.f1); if (lib.f2) fclose(lib.f2); }
	}

	mc_profile_stats pstat; memset(&pstat, 0, sizeof(pstat));
	if (ctx != NULL)
	{
		mc_totals t; mc_get_totals(ctx, &t);
		iTotalReadNum = t.total_reads; iTotalMappingNum = t.total_mapped; iTotalPairedNum = t.total_paired; TotalPairedDistance = t.total_distance; ReadLengthSum = t.read_length_sum;
		avgDist = t.avg_dist;
		if (bVCFoutput)
		{
			// device profile -> the structures VariantCalling() reads
			if (mc_profile_read(ctx, 0, GenomeSize, MappingRecordArr)) die("mc_profile_read");
			// CheckMappingCoverage / ReportDuplicationRate (src/ReadMapping.cpp:648-687) as device reductions
			if (mc_profile_summary(ctx, &pstat)) die("mc_profile_summary");
			iAlignedBase = pstat.aligned_bases; iTotalCoverage = pstat.coverage_sum;
			const mc_indel_rec* ir; int64_t ni; const uint8_t* arena;
			if (mc_profile_indels(ctx, &ir, &ni, &arena)) die("mc_profile_indels");
			for (int64_t i = 0; i < ni; i++) (ir[i].kind == 0 ? InsertSeqMap : DeleteSeqMap)[ir[i].pos][string((const char*)arena + ir[i].seq_off, ir[i].len)] = (uint16_t)ir[i].count;
			const mc_breakpoint_rec* br; int64_t nb;
			if (mc_profile_breakpoints(ctx, &br, &nb)) die("mc_profile_breakpoints");
			for (int64_t i = 0; i < nb; i++) BreakPointMap[br[i].pos] = (uint16_t)br[i].count;
			for (int kind = 0; kind < 2; kind++)
			{
				const mc_site_rec* sr; int64_t ns;
				if (mc_profile_sites(ctx, kind, &sr, &ns)) die("mc_profile_sites");
				vector<DiscordPair_t>& dst = kind == 0 ? InversionSiteVec : TranslocationSiteVec;
				for (int64_t i = 0; i < ns; i++) { DiscordPair_t d; d.gPos = sr[i].gPos; d.dist = sr[i].dist; dst.push_back(d); }
				sort(dst.begin(), dst.end(), CompByDiscordPos);   // the thread-end sort of the reference (src/ReadMapping.cpp:629-630)
			}
			// IdentifyVariants on the device, parked for the kept VariantCalling() (see the top of this file)
			mc_vc_params vp; mc_vc_params_default(&vp);
			vp.min_allele_depth = MinAlleleDepth; vp.frequency_thr = FrequencyThr; vp.somatic = bSomatic; vp.gvcf = bGVCF; vp.monomorphic = bMonomorphic;
			vp.ploidy = iPloidy; vp.min_cnv_size = MinCNVsize; vp.min_unmapped_size = MinUnmappedSize;
			const mc_variant_rec* vr; int64_t nv, nblk; const uint8_t* alt; const int32_t* depth;
			if (mc_variant_scan(ctx, &vp, &vr, &nv, &alt, &depth, &nblk)) die("mc_variant_scan");
			g_scanned_variants.resize((size_t)nv);
			for (int64_t i = 0; i < nv; i++)
			{
				Variant_t& v = g_scanned_variants[i];
				v.gPos = vr[i].gPos; v.DP = vr[i].DP; v.AD_ref = vr[i].AD_ref; v.AD_alt = vr[i].AD_alt; v.GenoType = vr[i].GenoType; v.qscore = vr[i].qscore; v.VarType = vr[i].VarType;
				if (vr[i].VarType == MC_VAR_INS || vr[i].VarType == MC_VAR_DEL) v.ALTstr.assign((const char*)alt + vr[i].alt_off, vr[i].alt_len);
				else if (vr[i].VarType == MC_VAR_SUB) v.ALTstr.assign(vr[i].alt, strnlen(vr[i].alt, 3));
			}
		}
		mc_ctx_destroy(ctx);
	}
	mc_index_free(idx);

	// run summary, as the reference prints it (src/ReadMapping.cpp:749-790)
	FILE* log = fopen(LogFileName, "a");
	const char* kind = bPairEnd ? "paired-end" : "single-end";
	const long long secs = (long long)(time(NULL) - StartProcessTime);
	fprintf(log, "All the %lld %s reads have been processed in %lld seconds.\n", (long long)iTotalReadNum, kind, secs);
	fprintf(stderr, "\rAll the %lld %s reads have been processed in %lld seconds.\n", (long long)iTotalReadNum, kind, secs);
	if (iTotalReadNum > 0)
	{
		double pct = (int)(10000 * (1.0 * iTotalMappingNum / iTotalReadNum) + 0.00005) / 100.0;
		fprintf(log, "%12lld (%6.2f%%) reads are mapped properly.\n", (long long)iTotalMappingNum, pct);
		fprintf(stderr, "%12lld (%6.2f%%) reads are mapped properly.\n", (long long)iTotalMappingNum, pct);
	}
	if (iTotalReadNum > 0 && iTotalPairedNum > 0)
	{
		double pct = (int)(10000 * (1.0 * (iTotalPairedNum << 1) / iTotalReadNum) + 0.00005) / 100.0;
		fprintf(log, "%12lld (%6.2f%%) reads are mapped in pairs.\n", (long long)(iTotalPairedNum << 1), pct);
		fprintf(stderr, "%12lld (%6.2f%%) reads are mapped in pairs.\n", (long long)(iTotalPairedNum << 1), pct);
	}
	if (bSAMoutput) fclose(sam_out);
	if (bVCFoutput)
	{
		avgCov = (int)(1.0 * iTotalCoverage / iAlignedBase + .5);
		fprintf(log, "\tEstimated AvgCoverage = %d\n", avgCov); fprintf(stderr, "\tEstimated AvgCoverage = %d\n", avgCov);
		const int64_t sites = pstat.dup_sites, total = pstat.dup_reads - pstat.dup_sites;
		fprintf(log, "\tDuplication rate=%4.2f%%\n", 100 * (1.0 * total / sites)); fprintf(stderr, "\tDuplication rate=%4.2f%%\n", 100 * (1.0 * total / sites));
	}
	if (iTotalReadNum > 0 && iTotalPairedNum > 0)
	{
		avgDist = (int)(1. * TotalPairedDistance / iTotalPairedNum + .5);
		avgReadLength = (int)(1. * ReadLengthSum / (iTotalPairedNum << 1) + .5);
		FragmentSize = avgDist + avgReadLength;
		fprintf(log, "\tAverage read length = %d, Estimated fragment size = %d, insert size = %d\n", avgReadLength, FragmentSize, avgDist - avgReadLength);
		fprintf(stderr, "\tAverage read length = %d, Estimated fragment size = %d, insert size = %d\n", avgReadLength, FragmentSize, avgDist - avgReadLength);
	}
	else avgDist = avgReadLength = 0;
	fclose(log);
}
