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
#include <unistd.h>
#include "../../../include/minialign_b200.h"
#include "mab_sam.h"
#include "mab_index.h"
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* ---- .mai loader ---- */
/* The container is a sequence of independently deflated frames (<= 1 MiB raw each): read them all, inflate them on all host
 * cores (the reference does the same with its worker threads), concatenate. */
static bool load_mai(const char *path, std::vector<uint8_t> &blob)
{
	FILE *fp = fopen(path, "rb");
	if(!fp) { return false; }
	std::vector<std::vector<uint8_t>> frames;
	while(true) {
		char magic[4]; uint32_t len;
		if(fread(magic, 1, 4, fp) != 4 || memcmp(magic, "PG00", 4) != 0) { break; }
		if(fread(&len, 4, 1, fp) != 1 || len == 0xffffffffu || len == 0) { break; }
		frames.emplace_back(len);
		if(fread(frames.back().data(), 1, len, fp) != len) { fclose(fp); return false; }
	}
	fclose(fp);
	if(frames.empty()) { return false; }
	std::vector<std::vector<uint8_t>> raws(frames.size());
	std::atomic<size_t> next(0); std::atomic<bool> ok(true);
	auto work = [&]() {
		std::vector<uint8_t> obuf(1 << 21);
		for(size_t i; (i = next.fetch_add(1)) < frames.size();) {
			z_stream zs; memset(&zs, 0, sizeof(zs));
			zs.next_in = frames[i].data(); zs.avail_in = (uInt)frames[i].size(); zs.next_out = obuf.data(); zs.avail_out = (uInt)obuf.size();
			if(inflateInit2(&zs, 15) != Z_OK) { ok = false; return; }
			int rc = inflate(&zs, Z_FINISH); inflateEnd(&zs);
			if(rc != Z_STREAM_END) { ok = false; return; }
			raws[i].assign(obuf.data(), obuf.data() + (obuf.size() - zs.avail_out));
		}
	};
	unsigned nth = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
	std::vector<std::thread> th;
	for(unsigned t = 1; t < nth; t++) { th.emplace_back(work); }
	work();
	for(auto &x : th) { x.join(); }
	if(!ok) { return false; }
	std::vector<uint8_t> raw;
	size_t tot = 0; for(auto &r : raws) { tot += r.size(); }
	raw.reserve(tot);
	for(auto &r : raws) { raw.insert(raw.end(), r.begin(), r.end()); }
	if(raw.size() < 12) { return false; }
	uint32_t magic; uint64_t size; memcpy(&magic, raw.data(), 4); memcpy(&size, raw.data() + 4, 8);
	if(magic != 0x0849414du || raw.size() < 12 + size) { return false; }
	blob.assign(raw.begin() + 12, raw.begin() + 12 + size);
	return true;
}

/* ---- FASTA / FASTQ reader (gz transparent) ---- */
/* Chunked: 4 MiB reads, memchr for the line ends, whole lines encoded through a 256-entry table (the low-nibble table encaf,
 * minialign.c:214-232) straight into the record. */
struct Rec { std::string name, qual; std::vector<uint8_t> seq; };
struct SeqReader {
	gzFile fp; std::vector<char> buf; size_t pos = 0, end = 0; bool eof = false;
	std::string line; bool have = false;
	uint8_t enc[256];
	explicit SeqReader(const char *path) : buf(4 << 20) {
		fp = strcmp(path, "-") == 0 ? gzdopen(0, "rb") : gzopen(path, "rb");
		if(fp) { gzbuffer(fp, 1 << 20); }
		uint8_t e16[16]; memset(e16, 0, sizeof(e16));
		e16['A' & 15] = 0; e16['C' & 15] = 1; e16['G' & 15] = 2; e16['T' & 15] = 3; e16['U' & 15] = 3; e16['N' & 15] = 4;
		for(int i = 0; i < 256; i++) { enc[i] = e16[i & 15]; }
	}
	~SeqReader() { if(fp) { gzclose(fp); } }
	bool fill() {
		if(eof) { return false; }
		if(pos < end) { memmove(buf.data(), buf.data() + pos, end - pos); }
		end -= pos; pos = 0;
		if(end == buf.size()) { buf.resize(buf.size() * 2); }
		int n = gzread(fp, buf.data() + end, (unsigned)(buf.size() - end));
		if(n <= 0) { eof = true; return false; }
		end += (size_t)n;
		return true;
	}
	/* next line without its terminator as [*b, *b + *n); false at end of input */
	bool getline(const char **b, size_t *n) {
		if(have) { have = false; *b = line.data(); *n = line.size(); return true; }
		while(true) {
			const char *nl = pos < end ? (const char *)memchr(buf.data() + pos, '\n', end - pos) : nullptr;
			if(nl) {
				*b = buf.data() + pos; *n = (size_t)(nl - *b); pos = (size_t)(nl - buf.data()) + 1;
				if(*n && (*b)[*n - 1] == '\r') { (*n)--; }
				return true;
			}
			if(!fill()) {
				if(pos < end) { *b = buf.data() + pos; *n = end - pos; pos = end; return true; }
				return false;
			}
		}
	}
	void unget(const char *b, size_t n) { line.assign(b, n); have = true; }
	bool next(Rec &r) {
		r.name.clear(); r.qual.clear(); r.seq.clear();
		const char *b; size_t n;
		while(getline(&b, &n)) {
			if(n == 0 || (b[0] != '>' && b[0] != '@')) { continue; }
			bool fq = b[0] == '@';
			size_t e = 1; while(e < n && b[e] != ' ' && b[e] != '\t') { e++; }
			r.name.assign(b + 1, e - 1);
			r.seq.reserve(32768);
			while(getline(&b, &n)) {
				if(!fq && n && b[0] == '>') { unget(b, n); break; }
				if(fq && n && b[0] == '+') { break; }
				size_t o = r.seq.size(); r.seq.resize(o + n);
				for(size_t i = 0; i < n; i++) { r.seq[o + i] = enc[(uint8_t)b[i]]; }
			}
			if(fq) { while(r.qual.size() < r.seq.size() && getline(&b, &n)) { r.qual.append(b, n); } }
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
	/* Pipeline: a reader thread parses and packs the next batches (reads already sit 1 byte/base with 64 B margins, the layout of
	 * bseq_t, minialign.c:2109-2146) while this thread maps the current one; the SAM text of a batch is formatted by all host
	 * cores, one slice of reads each, and written in input order.  One context, batches in input order: the reference's
	 * per-thread state (rlen) carries over exactly as with -t1. */
	struct Batch { std::vector<Rec> recs; std::vector<uint8_t> block; std::vector<uint64_t> ofs; std::vector<uint32_t> len; uint64_t bases = 0; size_t file = 0; bool last_of_file = false; bool open_failed = false; };
	std::mutex mu; std::condition_variable cv;
	std::deque<std::unique_ptr<Batch>> queue; bool done = false, stop = false;
	std::thread reader([&]() {
		for(size_t qi = 1; qi < o.pos.size(); qi++) {
			SeqReader rd(o.pos[qi].c_str());
			if(!rd.fp) { auto bt = std::make_unique<Batch>(); bt->file = qi; bt->open_failed = true; std::unique_lock<std::mutex> lk(mu); queue.push_back(std::move(bt)); cv.notify_all(); break; }
			bool more = true;
			while(more) {
				auto bt = std::make_unique<Batch>(); bt->file = qi;
				Rec r;
				while(bt->recs.size() < o.batch_reads && bt->bases < o.batch_bases && (more = rd.next(r))) { if(r.seq.empty()) { continue; } bt->bases += r.seq.size(); bt->recs.push_back(std::move(r)); }
				bt->last_of_file = !more;
				if(!bt->recs.empty()) {
					bt->block.assign(64 + bt->bases + 64ull * bt->recs.size() + 64, 0);
					bt->ofs.resize(bt->recs.size()); bt->len.resize(bt->recs.size());
					uint64_t p = 64;
					for(size_t i = 0; i < bt->recs.size(); i++) {
						bt->ofs[i] = p; bt->len[i] = (uint32_t)bt->recs[i].seq.size();
						memcpy(bt->block.data() + p, bt->recs[i].seq.data(), bt->len[i]); p += bt->len[i] + 64;
						std::vector<uint8_t>().swap(bt->recs[i].seq);
					}
				}
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&]() { return queue.size() < 2 || stop; });
				if(stop) { return; }
				queue.push_back(std::move(bt)); cv.notify_all();
			}
		}
		std::unique_lock<std::mutex> lk(mu); done = true; cv.notify_all();
	});
	double t_idx = now() - t0;
	mab_ctx *ctx = mab_init(blob.data(), blob.size(), &o.p, o.device);
	if(!ctx) {
		fprintf(stderr, "[E::main_align] failed to instanciate alignment context: %s\n", mab_last_error());
		{ std::unique_lock<std::mutex> lk(mu); stop = true; cv.notify_all(); }
		reader.join();
		return 1;
	}
	uint32_t n_ref = mab_n_ref(ctx);
	std::vector<MabSamRef> refs(n_ref);
	for(uint32_t i = 0; i < n_ref; i++) { mab_ref_info(ctx, i, &refs[i].name, &refs[i].l_name, &refs[i].l_seq, &refs[i].seq); }
	fprintf(stderr, "[M::main_align::%.3f] loaded/built index for %u target sequence(s) (index file %.3f s, device context %.3f s).\n", now() - t0, n_ref, t_idx, now() - t0 - t_idx);
	double tmap = now(); uint64_t tot_bases = 0, tot_reads = 0;
	std::string out; out.reserve(64 << 20);
	mab_sam_header(out, refs.data(), n_ref, "0.6.0-devel", cmdline.c_str());
	unsigned nfmt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
	/* writer stage: takes (batch, detached results) in input order, formats the SAM text on all cores, writes it */
	struct Done { std::unique_ptr<Batch> bt; mab_results *res; };
	std::mutex wmu; std::condition_variable wcv; std::deque<Done> wq; bool wdone = false;
	double t_fmt = 0, t_wr = 0;
	std::thread writer([&]() {
		std::vector<std::string> parts(nfmt);									/* per-slice SAM text, reused from batch to batch */
		if(!out.empty()) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }		/* the header */
		while(true) {
			Done d;
			{
				std::unique_lock<std::mutex> lk(wmu);
				wcv.wait(lk, [&]() { return !wq.empty() || wdone; });
				if(wq.empty()) { break; }
				d = std::move(wq.front()); wq.pop_front(); wcv.notify_all();
			}
			Batch *bt = d.bt.get();
			double tm0 = now();
			size_t n = bt->recs.size(), nsl = std::min<size_t>(nfmt, (n + 63) / 64);
			for(auto &pz : parts) { pz.clear(); }
			auto fmt = [&](size_t t) {
				std::string &dst = parts[t];
				size_t lo = n * t / nsl, hi = n * (t + 1) / nsl, est = 0;
				for(size_t i = lo; i < hi; i++) { est += bt->len[i] + bt->len[i] / 3 + 512; }
				if(dst.capacity() < est) { dst.reserve(est + est / 8); }
				for(size_t i = lo; i < hi; i++) {
					const uint32_t *w = nullptr; uint64_t nw = mab_results_get(d.res, (uint32_t)i, &w);
					MabSamRead q = { bt->recs[i].name.c_str(), (uint32_t)bt->recs[i].name.size(), bt->block.data() + bt->ofs[i], bt->len[i], bt->recs[i].qual.empty() ? nullptr : bt->recs[i].qual.c_str() };
					mab_sam_record(dst, refs.data(), &q, w, nw, o.tags);
				}
			};
			std::vector<std::thread> th;
			for(size_t t = 1; t < nsl; t++) { th.emplace_back(fmt, t); }
			if(nsl) { fmt(0); }
			for(auto &x : th) { x.join(); }
			mab_results_free(d.res);
			t_fmt += now() - tm0; tm0 = now();
			for(size_t t = 0; t < nsl; t++) { fwrite(parts[t].data(), 1, parts[t].size(), stdout); }
			t_wr += now() - tm0;
		}
	});
	int rc_main = 0;
	double t_wait = 0, t_map = 0, t_wwait = 0;
	while(true) {
		std::unique_ptr<Batch> bt;
		double tq = now();
		{
			std::unique_lock<std::mutex> lk(mu);
			cv.wait(lk, [&]() { return !queue.empty() || done; });
			if(queue.empty()) { break; }
			bt = std::move(queue.front()); queue.pop_front(); cv.notify_all();
		}
		t_wait += now() - tq;
		if(rc_main) { continue; }												/* drain the queue after an error */
		if(bt->open_failed) { fprintf(stderr, "[E::main_align] failed to open sequence file `%s'. Please check file path and format.\n", o.pos[bt->file].c_str()); rc_main = 1; continue; }
		bool last = bt->last_of_file; size_t file = bt->file;
		if(!bt->recs.empty()) {
			double tm0 = now();
			int rc = mab_map_batch(ctx, bt->block.data(), bt->block.size(), bt->ofs.data(), bt->len.data(), (uint32_t)bt->recs.size());
			t_map += now() - tm0;
			if(rc != MAB_OK) { fprintf(stderr, "[E::main_align] failed to map sequence file `%s': %s\n", o.pos[bt->file].c_str(), mab_last_error()); rc_main = 1; continue; }
			tot_bases += bt->bases; tot_reads += bt->recs.size();
			Done d; d.res = mab_detach_batch(ctx); d.bt = std::move(bt);
			tm0 = now();
			std::unique_lock<std::mutex> lk(wmu);
			wcv.wait(lk, [&]() { return wq.size() < 2; });
			wq.push_back(std::move(d)); wcv.notify_all();
			t_wwait += now() - tm0;
		}
		if(last) { fprintf(stderr, "[M::main_align::%.3f] finished mapping `%s' onto `%s'.\n", now() - t0, o.pos[file].c_str(), o.pos[0].c_str()); }
	}
	{ std::unique_lock<std::mutex> lk(wmu); wdone = true; wcv.notify_all(); }
	writer.join();
	reader.join();
	if(rc_main) { mab_destroy(ctx); return rc_main; }
	fprintf(stderr, "[M::main_align] host pipeline: mapper waited for the reader %.3f s and for the writer %.3f s, mapping %.3f s; writer: SAM formatting %.3f s, writing %.3f s\n", t_wait, t_wwait, t_map, t_fmt, t_wr);
	fwrite(out.data(), 1, out.size(), stdout);
	double tm = now() - tmap;
	fprintf(stderr, "[M::main] mapped %llu reads / %.1f Mbases in %.3f sec (%.1f Mbases/s)\n", (unsigned long long)tot_reads, tot_bases / 1e6, tm, tot_bases / 1e6 / tm);
	fprintf(stderr, "[M::main] Command: %s\n[M::main] Real time: %.3f sec\n", cmdline.c_str(), now() - t0);
	/* everything is written: leave without unpinning the host pools and freeing the DP arenas one by one (0.5-2 s for nothing;
	 * the driver reclaims the context with the process) */
	fflush(stdout); fflush(stderr);
	_exit(0);
}
