/*
 * mab_cli.cpp -- `minialign-b200`: minialign's command line for the mapping path (minialign.c:5703-6484), host side in C++.
 *
 *   minialign-b200 [-x preset] [-t N] [-T tag,tag] [-a -b -p -q -r -Y -s -m -W -G ...] ref.mai reads.fa[.gz] [...] > out.sam
 *
 *   minialign-b200 [-x preset] [-k -w -f -B] -d out.mai ref.fa                                      (index construction)
 *
 * Loads a prebuilt .mai index ("PG00" framed zlib stream, minialign.c:1135-1502, 3136-3167) or builds the index from a FASTA
 * reference on the host (mab_index.cpp), parses FASTA/FASTQ into the reference's 1 byte/base codes (minialign.c:214-232),
 * maps batches through the C ABI (libminialign_b200.so, CUDA, no CPU fallback) and prints SAM with mab_sam.cpp.
 */
#include "../../../include/minialign_b200.h"
#include "mab_sam.h"
#include "mab_index.h"
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* ---- .mai loader ---- */
static bool load_mai(const char *path, std::vector<uint8_t> &blob)
{
	FILE *fp = fopen(path, "rb");
	if(!fp) { return false; }
	std::vector<uint8_t> raw, cbuf, obuf(1 << 21);
	while(true) {
		char magic[4]; uint32_t len;
		if(fread(magic, 1, 4, fp) != 4 || memcmp(magic, "PG00", 4) != 0) { break; }
		if(fread(&len, 4, 1, fp) != 1 || len == 0xffffffffu || len == 0) { break; }
		cbuf.resize(len);
		if(fread(cbuf.data(), 1, len, fp) != len) { fclose(fp); return false; }
		z_stream zs; memset(&zs, 0, sizeof(zs));
		zs.next_in = cbuf.data(); zs.avail_in = len; zs.next_out = obuf.data(); zs.avail_out = (uInt)obuf.size();
		if(inflateInit2(&zs, 15) != Z_OK) { fclose(fp); return false; }
		int rc = inflate(&zs, Z_FINISH); inflateEnd(&zs);
		if(rc != Z_STREAM_END) { fclose(fp); return false; }
		raw.insert(raw.end(), obuf.data(), obuf.data() + (obuf.size() - zs.avail_out));
	}
	fclose(fp);
	if(raw.size() < 12) { return false; }
	uint32_t magic; uint64_t size; memcpy(&magic, raw.data(), 4); memcpy(&size, raw.data() + 4, 8);
	if(magic != 0x0849414du || raw.size() < 12 + size) { return false; }
	blob.assign(raw.begin() + 12, raw.begin() + 12 + size);
	return true;
}

/* ---- FASTA / FASTQ reader (gz transparent) ---- */
struct Rec { std::string name, qual; std::vector<uint8_t> seq; };
struct SeqReader {
	gzFile fp; std::string line; bool have = false, eof = false;
	uint8_t enc[16];
	explicit SeqReader(const char *path) {
		fp = strcmp(path, "-") == 0 ? gzdopen(0, "rb") : gzopen(path, "rb");
		memset(enc, 0, sizeof(enc));
		enc['A' & 15] = 0; enc['C' & 15] = 1; enc['G' & 15] = 2; enc['T' & 15] = 3; enc['U' & 15] = 3; enc['N' & 15] = 4;	/* encaf: low nibble table */
	}
	~SeqReader() { if(fp) { gzclose(fp); } }
	bool getline() {
		if(have) { have = false; return true; }
		line.clear();
		char buf[65536];
		while(true) {
			if(!gzgets(fp, buf, sizeof(buf))) { eof = true; return !line.empty(); }
			size_t l = strlen(buf); line.append(buf, l);
			if(l && buf[l - 1] == '\n') { break; }
		}
		while(!line.empty() && (line.back() == '\n' || line.back() == '\r')) { line.pop_back(); }
		return true;
	}
	bool next(Rec &r) {
		r.name.clear(); r.qual.clear(); r.seq.clear();
		while(getline()) {
			if(line.empty()) { continue; }
			if(line[0] != '>' && line[0] != '@') { continue; }
			bool fq = line[0] == '@';
			size_t e = 1; while(e < line.size() && line[e] != ' ' && line[e] != '\t') { e++; }
			r.name = line.substr(1, e - 1);
			while(getline()) {
				if(!fq && !line.empty() && line[0] == '>') { have = true; break; }
				if(fq && !line.empty() && line[0] == '+') { break; }
				for(char c : line) { r.seq.push_back(enc[(uint8_t)c & 15]); }
			}
			if(fq) { while(r.qual.size() < r.seq.size() && getline()) { r.qual += line; } }
			return true;
		}
		return false;
	}
};

struct Opts {
	mab_params_t p; uint32_t tags = 0; int device = 0; uint32_t batch_reads = 16384; uint64_t batch_bases = 400ull << 20;
	MabIdxParams ip; bool w_set = false; std::string dump;
	std::vector<std::string> pos;
};

static void set_match(mab_params_t &p, int m) { for(int i = 0; i < 16; i++) { if((i & 3) == (i >> 2)) { p.score_matrix[i] = (int8_t)m; } } }
static void set_mismatch(mab_params_t &p, int x) { for(int i = 0; i < 16; i++) { if((i & 3) != (i >> 2)) { p.score_matrix[i] = (int8_t)-x; } } }

static bool apply_opt(Opts &o, char c, const char *arg);
static void apply_line(Opts &o, const char *line)
{
	std::string s(line); size_t i = 0;
	while(i < s.size()) {
		while(i < s.size() && s[i] == ' ') { i++; }
		if(i + 1 >= s.size() || s[i] != '-') { break; }
		char c = s[i + 1]; size_t e = i + 2; while(e < s.size() && s[e] != ' ') { e++; }
		apply_opt(o, c, s.substr(i + 2, e - i - 2).c_str());
		i = e;
	}
}
/* preset tree of minialign.c:5853-5878 */
static bool apply_preset(Opts &o, const char *arg)
{
	std::string a(arg);
	std::vector<std::string> tok; size_t st = 0;
	for(size_t i = 0; i <= a.size(); i++) { if(i == a.size() || a[i] == '.' || a[i] == ':') { tok.push_back(a.substr(st, i - st)); st = i + 1; } }
	if(tok.empty()) { return false; }
	if(tok[0] == "pacbio") { apply_line(o, "-k15 -w10 -a2 -b4 -p4 -q2 -r3,3 -Y50 -s50 -m0.3"); if(tok.size() > 1 && tok[1] == "ccs") { apply_line(o, "-b5 -p6 -p2"); } return true; }
	if(tok[0] == "ava") { apply_line(o, "-k15 -w5 -a2 -b3 -p0 -q2 -Y50 -s30 -m0.05"); return true; }
	if(tok[0] == "ont") {
		apply_line(o, "-k15 -w10 -a3 -b5 -p6 -q2 -r3,3 -Y50 -s50 -m0.3");
		for(size_t i = 1; i < tok.size(); i++) {
			const std::string &t = tok[i];
			if(t == "r7") { apply_line(o, "-b4"); }
			else if(t == "4" || t == "5") { apply_line(o, "-a2"); }
			else if(t == "1dsq" || t == "2d") { if(i == 1) { apply_line(o, "-a2"); } if(!(i >= 2 && tok[1] == "r7")) { apply_line(o, "-b6 -r4,4"); } }
			else if(t == "1d") { if(i == 1) { apply_line(o, "-a2"); } }
		}
		return true;
	}
	return false;
}
static bool apply_opt(Opts &o, char c, const char *arg)
{
	switch(c) {
		case 'x': return apply_preset(o, arg);
		case 'a': set_match(o.p, atoi(arg)); return true;
		case 'b': set_mismatch(o.p, atoi(arg)); return true;
		case 'p': o.p.gi = (int8_t)atoi(arg); return true;
		case 'q': o.p.ge = (int8_t)atoi(arg); return true;
		case 'r': { int a = atoi(arg); const char *cm = strchr(arg, ','); int b = cm ? atoi(cm + 1) : a; o.p.gfa = (int8_t)a; o.p.gfb = (int8_t)b; return true; }
		case 'Y': o.p.xdrop = (int8_t)atoi(arg); return true;
		case 's': o.p.min_score = (uint32_t)atoi(arg); return true;
		case 'm': o.p.min_ratio = (float)atof(arg); return true;
		case 'W': o.p.wlen = atoi(arg); return true;
		case 'G': o.p.glen = atoi(arg); return true;
		case 'T': o.tags |= mab_sam_parse_tags(arg); return true;
		case 't': return true;									/* host worker threads: the mapping runs on the GPU */
		case 'k': o.ip.k = (uint32_t)atoi(arg); return true;	/* index-time parameters: used when the index is built here, a .mai carries its own */
		case 'w': o.ip.w = (uint32_t)atoi(arg); o.w_set = true; return true;
		case 'B': o.ip.b = (uint32_t)atoi(arg); return true;
		case 'f': {												/* minialign.c:6012-6020: descending frequency thresholds */
			o.ip.n_frq = 0;
			for(const char *q = arg; *q && o.ip.n_frq < 7;) { o.ip.frq[o.ip.n_frq++] = (float)atof(q); const char *cm = strchr(q, ','); if(!cm) { break; } q = cm + 1; }
			return o.ip.n_frq > 0;
		}
		case 'd': o.dump = arg; return true;
		case 'g': o.device = atoi(arg); return true;
		case 'n': o.batch_reads = (uint32_t)atoi(arg); return true;
		default: return false;
	}
}

int main(int argc, char **argv)
{
	double t0 = now();
	Opts o; memset(&o.p, 0, sizeof(o.p));
	o.p.wlen = 7000; o.p.glen = 7000; o.p.min_score = 50; o.p.min_ratio = 0.3f; o.p.xdrop = 50;	/* defaults, minialign.c:6152-6158 */
	set_match(o.p, 1); set_mismatch(o.p, 1); o.p.gi = 1; o.p.ge = 1;
	std::string cmdline;
	for(int i = 0; i < argc; i++) { if(i) { cmdline += ' '; } cmdline += argv[i]; }
	for(int i = 1; i < argc; i++) {
		const char *a = argv[i];
		if(a[0] == '-' && a[1] != '\0') {
			if(a[1] == 'v') { fprintf(stderr, "[M::main] Version: 0.6.0-devel, Build: B200 (sm_100a)\n"); return 0; }
			const char *arg = a[2] ? a + 2 : (i + 1 < argc ? argv[++i] : "");
			if(!apply_opt(o, a[1], arg)) { fprintf(stderr, "[E::main] unknown or unsupported option `-%c'.\n", a[1]); return 1; }
		} else { o.pos.push_back(a); }
	}
	if(!o.w_set) { o.ip.w = (uint32_t)(int)(2.0 / 3.0 * o.ip.k + .499); }							/* default window size when -w is absent (minialign.c:6111) */
	if(o.pos.size() < (o.dump.empty() ? 2u : 1u)) { fprintf(stderr, "usage: minialign-b200 [-x preset] [-T tags] <ref.fa|ref.mai> <reads.fa> [...] > out.sam\n       minialign-b200 [-x preset] -d <out.mai> <ref.fa>\n"); return 1; }
	std::vector<uint8_t> blob;
	if(!load_mai(o.pos[0].c_str(), blob)) {															/* not an index: a FASTA reference, build it here */
		SeqReader rr(o.pos[0].c_str());
		if(!rr.fp) { fprintf(stderr, "[E::main_align] failed to open index file `%s'. Please check file path and it exists.\n", o.pos[0].c_str()); return 1; }
		std::vector<MabIdxSeq> refs; Rec r;
		while(rr.next(r)) { if(r.seq.empty()) { continue; } MabIdxSeq q; q.name = r.name; q.seq = std::move(r.seq); refs.push_back(std::move(q)); }
		std::string err;
		if(!mab_build_index(refs, o.ip, blob, err)) { fprintf(stderr, "[E::main_index] failed to build index from `%s': %s\n", o.pos[0].c_str(), err.c_str()); return 1; }
		fprintf(stderr, "[M::main_index::%.3f] built index for %zu target sequence(s).\n", now() - t0, refs.size());
	}
	if(!o.dump.empty()) {
		if(!mab_write_mai(o.dump.c_str(), blob)) { fprintf(stderr, "[E::main_index] failed to write index to `%s'.\n", o.dump.c_str()); return 1; }
		fprintf(stderr, "[M::main] Command: %s\n[M::main] Real time: %.3f sec\n", cmdline.c_str(), now() - t0);
		return 0;
	}
	mab_ctx *ctx = mab_init(blob.data(), blob.size(), &o.p, o.device);
	if(!ctx) { fprintf(stderr, "[E::main_align] failed to instanciate alignment context: %s\n", mab_last_error()); return 1; }
	uint32_t n_ref = mab_n_ref(ctx);
	std::vector<MabSamRef> refs(n_ref);
	for(uint32_t i = 0; i < n_ref; i++) { mab_ref_info(ctx, i, &refs[i].name, &refs[i].l_name, &refs[i].l_seq, &refs[i].seq); }
	fprintf(stderr, "[M::main_align::%.3f] loaded/built index for %u target sequence(s).\n", now() - t0, n_ref);
	double tmap = now(); uint64_t tot_bases = 0, tot_reads = 0;
	std::string out; out.reserve(64 << 20);
	mab_sam_header(out, refs.data(), n_ref, "0.6.0-devel", cmdline.c_str());
	for(size_t qi = 1; qi < o.pos.size(); qi++) {
		SeqReader rd(o.pos[qi].c_str());
		if(!rd.fp) { fprintf(stderr, "[E::main_align] failed to open sequence file `%s'. Please check file path and format.\n", o.pos[qi].c_str()); return 1; }
		bool more = true;
		while(more) {
			std::vector<Rec> recs; uint64_t bases = 0;
			Rec r;
			while(recs.size() < o.batch_reads && bases < o.batch_bases && (more = rd.next(r))) { if(r.seq.empty()) { continue; } bases += r.seq.size(); recs.push_back(std::move(r)); }
			if(recs.empty()) { break; }
			std::vector<uint8_t> block(64 + bases + 64ull * recs.size() + 64, 0);
			std::vector<uint64_t> ofs(recs.size()); std::vector<uint32_t> len(recs.size());
			uint64_t p = 64;
			for(size_t i = 0; i < recs.size(); i++) { ofs[i] = p; len[i] = (uint32_t)recs[i].seq.size(); memcpy(block.data() + p, recs[i].seq.data(), len[i]); p += len[i] + 64; }
			int rc = mab_map_batch(ctx, block.data(), block.size(), ofs.data(), len.data(), (uint32_t)recs.size());
			if(rc != MAB_OK) { fprintf(stderr, "[E::main_align] failed to map sequence file `%s': %s\n", o.pos[qi].c_str(), mab_last_error()); return 1; }
			for(size_t i = 0; i < recs.size(); i++) {
				const uint32_t *w = nullptr; uint64_t n = mab_result(ctx, (uint32_t)i, &w);
				MabSamRead q = { recs[i].name.c_str(), (uint32_t)recs[i].name.size(), block.data() + ofs[i], len[i], recs[i].qual.empty() ? nullptr : recs[i].qual.c_str() };
				mab_sam_record(out, refs.data(), &q, w, n, o.tags);
				if(out.size() > (48u << 20)) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
			}
			mab_release_batch(ctx);
			tot_bases += bases; tot_reads += recs.size();
		}
		fprintf(stderr, "[M::main_align::%.3f] finished mapping `%s' onto `%s'.\n", now() - t0, o.pos[qi].c_str(), o.pos[0].c_str());
	}
	fwrite(out.data(), 1, out.size(), stdout);
	double tm = now() - tmap;
	fprintf(stderr, "[M::main] mapped %llu reads / %.1f Mbases in %.3f sec (%.1f Mbases/s)\n", (unsigned long long)tot_reads, tot_bases / 1e6, tm, tot_bases / 1e6 / tm);
	fprintf(stderr, "[M::main] Command: %s\n[M::main] Real time: %.3f sec\n", cmdline.c_str(), now() - t0);
	mab_destroy(ctx);
	return 0;
}
