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
#include <map>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include <cerrno>
#include <memory>
#include <mutex>
#include <thread>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* ---- .mai loader ---- */
/* The container is a sequence of independently deflated frames of 1 MiB raw each (PG_BLOCK_SIZE, minialign.c:1137; the last one
 * shorter).  The file is mapped, the frame table read off it, and every frame is inflated by one of the host threads straight
 * into its place in the image: no intermediate copies (a human-sized index is ~15 GB raw). */
struct MaiImage {
	uint8_t *raw = nullptr; size_t raw_size = 0;		/* the inflated stream: 12-byte {magic, size} header + payload */
	const uint8_t *blob = nullptr; size_t size = 0;	/* the payload: what mab_init takes */
	std::vector<uint8_t> own;							/* payload built here (FASTA reference) */
	~MaiImage() { free(raw); }
};
/* on_size(payload bytes) is called once the header frame is read, on_piece(payload offset, bytes, n) for every inflated frame
 * from the thread that inflated it: the caller forwards the pieces to the GPUs while the other frames are still in the works */
struct MaiHooks { std::function<void(uint64_t)> on_size; std::function<void(uint64_t, const uint8_t *, uint64_t)> on_piece; };
static bool load_mai(const char *path, MaiImage &im, const MaiHooks *hooks = nullptr)
{
	int fd = open(path, O_RDONLY);
	if(fd < 0) { return false; }
	struct stat st;
	if(fstat(fd, &st) != 0 || st.st_size < 8) { close(fd); return false; }
	const size_t fsz = (size_t)st.st_size;
	const uint8_t *file = (const uint8_t *)mmap(nullptr, fsz, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if(file == MAP_FAILED) { return false; }
	struct Frame { size_t ofs; uint32_t len; };
	std::vector<Frame> frames;
	for(size_t p = 0; p + 8 <= fsz;) {
		uint32_t len; memcpy(&len, file + p + 4, 4);
		if(memcmp(file + p, "PG00", 4) != 0 || len == 0xffffffffu || len == 0 || p + 8 + len > fsz) { break; }
		frames.push_back({ p + 8, len }); p += 8 + (size_t)len;
	}
	if(frames.empty()) { munmap((void *)file, fsz); return false; }
	const size_t BS = 1 << 20;
	im.raw = (uint8_t *)malloc(frames.size() * BS + 64);
	if(!im.raw) { munmap((void *)file, fsz); return false; }
	std::vector<uint32_t> out_len(frames.size(), 0);
	std::atomic<size_t> next(1); std::atomic<bool> ok(true);
	uint64_t payload = 0;																		/* from the header in frame 0 */
	auto one = [&](size_t i) -> bool {
		z_stream zs; memset(&zs, 0, sizeof(zs));
		zs.next_in = (Bytef *)(file + frames[i].ofs); zs.avail_in = (uInt)frames[i].len; zs.next_out = im.raw + i * BS; zs.avail_out = (uInt)BS;
		if(inflateInit2(&zs, 15) != Z_OK) { return false; }
		int rc = inflate(&zs, Z_FINISH); inflateEnd(&zs);
		if(rc != Z_STREAM_END) { return false; }
		out_len[i] = (uint32_t)(BS - zs.avail_out);
		if(hooks && payload) {																	/* this frame's share of the payload [0, payload) */
			uint64_t lo = i * BS, hi = lo + out_len[i];
			lo = std::max<uint64_t>(lo, 12); hi = std::min<uint64_t>(hi, 12 + payload);
			if(lo < hi) { hooks->on_piece(lo - 12, im.raw + lo, hi - lo); }
		}
		return true;
	};
	{	/* frame 0 first: it holds the {magic, size} header the device buffers are sized by */
		z_stream zs; memset(&zs, 0, sizeof(zs));
		zs.next_in = (Bytef *)(file + frames[0].ofs); zs.avail_in = (uInt)frames[0].len; zs.next_out = im.raw; zs.avail_out = (uInt)BS;
		bool good = inflateInit2(&zs, 15) == Z_OK;
		if(good) { int rc = inflate(&zs, Z_FINISH); inflateEnd(&zs); good = rc == Z_STREAM_END; }
		out_len[0] = good ? (uint32_t)(BS - zs.avail_out) : 0;
		uint32_t magic = 0; uint64_t size = 0;
		if(good && out_len[0] >= 12) { memcpy(&magic, im.raw, 4); memcpy(&size, im.raw + 4, 8); }
		if(!good || magic != 0x0849414du || size + 12 > frames.size() * BS) { munmap((void *)file, fsz); return false; }
		if(hooks) {
			payload = size; hooks->on_size(size);
			uint64_t hi = std::min<uint64_t>(out_len[0], 12 + payload);
			if(hi > 12) { hooks->on_piece(0, im.raw + 12, hi - 12); }
		}
	}
	auto work = [&]() {
		for(size_t i; (i = next.fetch_add(1)) < frames.size();) { if(!one(i)) { ok = false; return; } }
	};
	unsigned nth = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
	if(hooks && nth > 2) { nth--; }															/* a core for the thread that brings the devices up meanwhile */
	std::vector<std::thread> th;
	for(unsigned t = 1; t < nth; t++) { th.emplace_back(work); }
	work();
	for(auto &x : th) { x.join(); }
	munmap((void *)file, fsz);
	if(!ok) { return false; }
	for(size_t i = 0; i + 1 < frames.size(); i++) { if(out_len[i] != BS) { return false; } }		/* not the layout the reference writes */
	im.raw_size = (frames.size() - 1) * BS + out_len.back();
	if(im.raw_size < 12) { return false; }
	uint32_t magic; uint64_t size; memcpy(&magic, im.raw, 4); memcpy(&size, im.raw + 4, 8);
	if(magic != 0x0849414du || im.raw_size < 12 + size) { return false; }
	im.blob = im.raw + 12; im.size = (size_t)size;
	return true;
}

/* ---- FASTA / FASTQ reader (gz transparent) ---- */
/* Chunked: 4 MiB reads, memchr for the line ends, whole lines encoded through a 256-entry table (the low-nibble table encaf,
 * minialign.c:214-232) straight into the record. */
struct Rec { std::string name, qual; std::vector<uint8_t> seq; };
struct SeqReader {
	gzFile fp; std::vector<char> buf; size_t pos = 0, end = 0; bool eof = false;
	std::string line; bool have = false;
	bool keep_qual = false;			/* -Q (minialign.c:5964, 2064): without it the quality lines are skipped */
	uint8_t enc[256];
	/* records out of a chunk of text already in memory */
	SeqReader(const char *mem, size_t n) : buf(mem, mem + n) { fp = nullptr; end = n; eof = true; init_enc(); }
	explicit SeqReader(const char *path) : buf(4 << 20) {
		fp = strcmp(path, "-") == 0 ? gzdopen(0, "rb") : gzopen(path, "rb");
		if(fp) { gzbuffer(fp, 1 << 20); }
		init_enc();
	}
	void init_enc() {
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
			if(fq) { std::string q; while(q.size() < r.seq.size() && getline(&b, &n)) { q.append(b, n); } if(keep_qual) { r.qual.swap(q); } }
			return true;
		}
		return false;
	}
};

/* ---- raw bytes of a read file (gzip transparent), for the text path ---- */
struct ByteSource {
	int fd = -1; gzFile gz = nullptr;
	explicit ByteSource(const char *path) {
		if(strcmp(path, "-") == 0) { gz = gzdopen(0, "rb"); if(gz) { gzbuffer(gz, 1 << 20); } return; }
		fd = open(path, O_RDONLY);
		if(fd < 0) { return; }
		unsigned char m[2] = { 0, 0 };
		ssize_t n = pread(fd, m, 2, 0);
		if(n == 2 && m[0] == 0x1f && m[1] == 0x8b) { close(fd); fd = -1; gz = gzopen(path, "rb"); if(gz) { gzbuffer(gz, 1 << 20); } }
	}
	~ByteSource() { if(fd >= 0) { close(fd); } if(gz) { gzclose(gz); } }
	bool ok() const { return fd >= 0 || gz != nullptr; }
	int64_t read(char *dst, uint64_t n) {
		if(n > (1u << 30)) { n = 1u << 30; }
		if(gz) { return gzread(gz, dst, (unsigned)n); }
		return ::read(fd, dst, n);
	}
};

/* [0, cut) = whole records of buf[0, fill): the start of the last record header that can be told for sure.  FASTA: the last
 * '>' at a line start.  FASTQ: the last '@' at a line start whose line after next starts with '+' (a quality line may start
 * with '@' too, but then the line after next is a sequence).  0 = no such place. */
static uint64_t record_cut(const char *buf, uint64_t fill, char delim = 0)
{
	if(fill < 2) { return 0; }
	if(delim == 0) { delim = buf[0]; }
	for(uint64_t p = fill - 1; p > 0; p--) {
		if(buf[p] != delim || buf[p - 1] != '\n') { continue; }
		if(delim != '@') { return p; }
		const char *e1 = (const char *)memchr(buf + p, '\n', fill - p);
		if(!e1) { continue; }
		const char *e2 = (const char *)memchr(e1 + 1, '\n', fill - (uint64_t)(e1 + 1 - buf));
		if(!e2 || (uint64_t)(e2 + 1 - buf) >= fill) { continue; }
		if(e2[1] == '+') { return p; }
	}
	return 0;
}

static int write_all(int fd, const char *p, uint64_t n)
{
	while(n) { ssize_t w = write(fd, p, n > (1u << 30) ? (1u << 30) : n); if(w < 0) { if(errno == EINTR) { continue; } return -1; } p += w; n -= (uint64_t)w; }
	return 0;
}

struct Opts {
	mab_params_t p; uint32_t tags = 0; int device = 0; uint32_t batch_reads = 16384; uint64_t batch_bases = 400ull << 20;
	bool keep_qual = false; unsigned contexts = 3; std::vector<int> devices; double chunk_mb = 320;
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
		case 'g': {												/* device ordinal, or a comma-separated list: the chunks are dealt to all of them */
			o.devices.clear();
			for(const char *q = arg; *q;) { o.devices.push_back(atoi(q)); const char *cm = strchr(q, ','); if(!cm) { break; } q = cm + 1; }
			if(o.devices.empty()) { return false; }
			o.device = o.devices[0];
			return true;
		}
		case 'Q': o.keep_qual = true; return true;				/* keep the FASTQ qualities (minialign.c:5964) */
		case 'c': o.contexts = (unsigned)std::max(1, atoi(arg)); return true;	/* chunks in flight per GPU */
		case 'N': o.chunk_mb = std::max(0.001, atof(arg)); return true;	/* chunk size in MiB of read file */
		case 'n': o.batch_reads = (uint32_t)atoi(arg); return true;
		default: return false;
	}
}

/* a chunk the device reader does not take: host reader -> record-level mapper -> host formatter (runs in file order, the
 * context's rlen word is set by hand) */
static bool map_chunk_on_host(mab_ctx *ctx, const char *buf, uint64_t len, const Opts &o, const std::vector<MabSamRef> &refs, uint32_t rlen_in, uint32_t *rlen_out,
	std::string &out, uint64_t *n_reads, uint64_t *n_bases)
{
	SeqReader rd(buf, len); rd.keep_qual = o.keep_qual;
	std::vector<Rec> recs; Rec r;
	uint64_t bases = 0;
	while(rd.next(r)) { if(r.seq.empty()) { continue; } bases += r.seq.size(); recs.push_back(std::move(r)); }
	*n_reads = recs.size(); *n_bases = bases; *rlen_out = rlen_in;
	if(recs.empty()) { return true; }
	std::vector<uint8_t> block(64 + bases + 64ull * recs.size() + 64, 0);
	std::vector<uint64_t> ofs(recs.size()); std::vector<uint32_t> lens(recs.size());
	uint64_t p = 64;
	for(size_t i = 0; i < recs.size(); i++) { ofs[i] = p; lens[i] = (uint32_t)recs[i].seq.size(); memcpy(block.data() + p, recs[i].seq.data(), lens[i]); p += lens[i] + 64; }
	mab_set_rlen(ctx, rlen_in);
	if(mab_map_batch(ctx, block.data(), block.size(), ofs.data(), lens.data(), (uint32_t)recs.size()) != MAB_OK) { return false; }
	*rlen_out = mab_get_rlen(ctx);
	for(size_t i = 0; i < recs.size(); i++) {
		const uint32_t *w = nullptr; uint64_t nw = mab_result(ctx, (uint32_t)i, &w);
		MabSamRead q = { recs[i].name.c_str(), (uint32_t)recs[i].name.size(), block.data() + ofs[i], lens[i], recs[i].qual.empty() ? nullptr : recs[i].qual.c_str() };
		mab_sam_record(out, refs.data(), &q, w, nw, o.tags);
	}
	mab_release_batch(ctx);
	return true;
}

int main(int argc, char **argv)
{
	double t0 = now();
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);				/* hardware queues per device: the contexts' streams (and their side streams) share 8 by default and then wait for each other's kernels */
	Opts o; memset(&o.p, 0, sizeof(o.p));
	o.p.wlen = 7000; o.p.glen = 7000; o.p.min_score = 50; o.p.min_ratio = 0.3f; o.p.xdrop = 50;	/* defaults, minialign.c:6152-6158 */
	set_match(o.p, 1); set_mismatch(o.p, 1); o.p.gi = 1; o.p.ge = 1;
	std::string cmdline;
	for(int i = 0; i < argc; i++) { if(i) { cmdline += ' '; } cmdline += argv[i]; }
	for(int i = 1; i < argc; i++) {
		const char *a = argv[i];
		if(a[0] == '-' && a[1] != '\0') {
			if(a[1] == 'v') { fprintf(stderr, "[M::main] Version: 0.6.0-devel, Build: B200 (sm_100a)\n"); return 0; }
			const char *arg = a[1] == 'Q' ? "" : (a[2] ? a + 2 : (i + 1 < argc ? argv[++i] : ""));
			if(!apply_opt(o, a[1], arg)) { fprintf(stderr, "[E::main] unknown or unsupported option `-%c'.\n", a[1]); return 1; }
		} else { o.pos.push_back(a); }
	}
	if(!o.w_set) { o.ip.w = (uint32_t)(int)(2.0 / 3.0 * o.ip.k + .499); }							/* default window size when -w is absent (minialign.c:6111) */
	if(o.pos.size() < (o.dump.empty() ? 2u : 1u)) { fprintf(stderr, "usage: minialign-b200 [-x preset] [-T tags] <ref.fa|ref.mai> <reads.fa> [...] > out.sam\n       minialign-b200 [-x preset] -d <out.mai> <ref.fa>\n"); return 1; }
	/* the host side of the pipeline below: page-locked chunk and output buffers.  Pinning memory is slow (1.5-2.5 GB/s), so a thread
	 * of its own does it while the first chunks are already being read and mapped; readers and workers pick the buffers up as they appear. */
	const uint64_t chunk_bytes = std::max<uint64_t>(1024, (uint64_t)(o.chunk_mb * 1048576.0));
	std::vector<int> devices = o.devices.empty() ? std::vector<int>{ o.device } : o.devices;
	const unsigned n_ctx = std::max(1u, o.contexts) * (unsigned)devices.size();
	struct Chunk { char *buf = nullptr; uint64_t cap = 0, len = 0; uint64_t id = 0; size_t file = 0; bool last_of_file = false, open_failed = false; };
	struct Out { char *buf = nullptr; uint64_t cap = 0, len = 0; std::string spill; uint64_t id = 0; unsigned owner = 0; bool busy = false; };	/* buf: page-locked, the SAM text is copied from the device straight into it */
	std::mutex mu; std::condition_variable cv;
	std::deque<Chunk *> free_chunks, ready; std::map<uint64_t, Out *> done_outs;
	bool read_done = false, failed = false, pin_stop = false, pin_go = false;
	std::vector<Chunk> chunk_pool(n_ctx + 2); std::vector<Out> out_pool(2 * n_ctx);			/* two output buffers per context: one being written while the next is filled */
	for(unsigned i = 0; i < 2 * n_ctx; i++) { out_pool[i].owner = i / 2; }
	std::thread pinner;
	if(o.dump.empty()) {
		pinner = std::thread([&]() {
			auto stopped = [&]() { std::unique_lock<std::mutex> lk(mu); return pin_stop; };
			auto give_up = [&]() { fprintf(stderr, "[E::main_align] %s\n", mab_last_error()); std::unique_lock<std::mutex> lk(mu); failed = true; cv.notify_all(); };
			/* Page-locking in two steps.  Now, next to the index load: map the buffers and touch every page on a few threads (no
			 * CUDA involved).  Once the contexts stand: register them with the driver one by one in the order they are needed
			 * (0.15 s per GB, and other CUDA calls get through meanwhile).  cudaHostAlloc does both in one call at 0.4-0.6 s per GB
			 * and holds up every other CUDA call of the process for its duration: measured on the CLI that delayed the device
			 * set-up from 1.3 to 2.4-4.3 s when done early and the first chunk by 1.7 s when done late. */
			const uint64_t ccap = chunk_bytes + (16 << 20), ocap = chunk_bytes + chunk_bytes / 2 + (1 << 20);
			const size_t nc = chunk_pool.size(), no = out_pool.size();
			const uint64_t total = (ccap + 4096) * nc + (ocap + 4096) * no;
			char *arena = (char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
			if(arena == (char *)MAP_FAILED) { arena = nullptr; }
			std::vector<char *> cb(nc, nullptr), ob(no, nullptr);
			if(arena) {
				char *p = arena;
				for(size_t i = 0; i < nc; i++) { cb[i] = p; p += (ccap + 4095) & ~4095ull; }
				for(size_t i = 0; i < no; i++) { ob[i] = p; p += (ocap + 4095) & ~4095ull; }
				std::atomic<uint64_t> nextpg(0); const uint64_t step = 64ull << 20;
				auto touch = [&]() { for(uint64_t a; (a = nextpg.fetch_add(step)) < total && !stopped();) { memset(arena + a, 0, std::min<uint64_t>(step, total - a)); } };
				std::vector<std::thread> th; for(int t = 0; t < 3; t++) { th.emplace_back(touch); }
				touch(); for(auto &x : th) { x.join(); }
			}
			if(getenv("MAB_TRACE")) { fprintf(stderr, "[pinner %.3f] pages touched\n", now() - t0); }
			{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return pin_go || pin_stop; }); }
			if(getenv("MAB_TRACE")) { fprintf(stderr, "[pinner %.3f] start\n", now() - t0); }
			auto lock = [&](char *pre, uint64_t cap) -> char * {
				if(pre && mab_host_register(devices[0], pre, cap) == MAB_OK) { return pre; }
				return (char *)mab_host_alloc_on(devices[0], cap);
			};
			/* first what the first chunks need (a chunk buffer and an output buffer per context), then the second halves */
			for(unsigned pass = 0; pass < 2; pass++) {
				for(size_t i = pass ? std::min<size_t>(n_ctx, nc) : 0; i < (pass ? nc : std::min<size_t>(n_ctx, nc)); i++) {
					if(stopped()) { return; }
					Chunk &c = chunk_pool[i];
					char *b = lock(cb[i], ccap);
					if(!b) { give_up(); return; }
					std::unique_lock<std::mutex> lk(mu); c.cap = ccap; c.buf = b; free_chunks.push_back(&c); cv.notify_all();
				}
				if(getenv("MAB_TRACE")) { fprintf(stderr, "[pinner %.3f] pass %u: chunk buffers done\n", now() - t0, pass); }
				for(unsigned i = pass; i < 2 * n_ctx; i += 2) {
					if(stopped()) { return; }
					char *b = lock(ob[i], ocap);
					if(!b) { give_up(); return; }
					std::unique_lock<std::mutex> lk(mu); out_pool[i].cap = ocap; out_pool[i].buf = b; cv.notify_all();
				}
			}
			if(getenv("MAB_TRACE")) { fprintf(stderr, "[pinner %.3f] done\n", now() - t0); }
		});
	}
	auto stop_pinner = [&]() { { std::unique_lock<std::mutex> lk(mu); pin_stop = true; cv.notify_all(); } if(pinner.joinable()) { pinner.join(); } };
	MaiImage im;
	/* a .mai index goes to the GPUs frame by frame while the other frames are still being inflated (mab_load_*).  The devices are
	 * set up (CUDA context, buffers: a second or more) by a thread of their own next to the inflation; frames finished before they
	 * are ready wait in a list, which every inflating thread helps to empty afterwards. */
	mab_loader *ld = nullptr; std::atomic<bool> ld_failed(false);
	mab_params_t ld_prm = o.p; ld_prm.flags |= MAB_FLAG_BORROW_INDEX;								/* `im` outlives the contexts */
	struct Piece { uint64_t off; const uint8_t *p; uint64_t n; };
	std::mutex ld_mu; std::deque<Piece> backlog; bool ld_ready = false; std::thread ld_thread;
	auto drain = [&]() {
		for(;;) {
			Piece q;
			{ std::unique_lock<std::mutex> lk(ld_mu); if(!ld_ready || backlog.empty()) { return; } q = backlog.front(); backlog.pop_front(); }
			if(ld && !ld_failed && mab_load_put(ld, q.off, q.p, q.n) != MAB_OK) { ld_failed = true; }
		}
	};
	MaiHooks hooks;
	hooks.on_size = [&](uint64_t size) {
		ld_thread = std::thread([&, size]() {
			mab_loader *l = mab_load_begin(size, &ld_prm, devices.data(), (int)devices.size());
			if(getenv("MAB_TRACE")) { fprintf(stderr, "[loader %.3f] devices ready\n", now() - t0); }
			{ std::unique_lock<std::mutex> lk(ld_mu); ld = l; if(!l) { ld_failed = true; } ld_ready = true; }
			drain();
		});
	};
	hooks.on_piece = [&](uint64_t off, const uint8_t *p, uint64_t n) {
		{ std::unique_lock<std::mutex> lk(ld_mu); backlog.push_back({ off, p, n }); }
		drain();
	};
	const bool is_mai = load_mai(o.pos[0].c_str(), im, o.dump.empty() && getenv("MAB_NO_STAGED_LOAD") == nullptr ? &hooks : nullptr);
	if(getenv("MAB_TRACE")) { fprintf(stderr, "[main %.3f] index file read\n", now() - t0); }
	if(ld_thread.joinable()) {
		ld_thread.join();
		size_t left; { std::unique_lock<std::mutex> lk(ld_mu); left = backlog.size(); }
		if(left > 64) {																			/* the devices came up late: the rest of the list on all cores */
			std::vector<std::thread> th;
			for(unsigned t = 1; t < std::max(1u, std::min(std::thread::hardware_concurrency(), 32u)); t++) { th.emplace_back(drain); }
			drain();
			for(auto &x : th) { x.join(); }
		} else { drain(); }
	}
	if(ld && (!is_mai || ld_failed)) { mab_load_abort(ld); ld = nullptr; }								/* the plain path below reports what is wrong */
	if(!is_mai) {															/* not an index: a FASTA reference, build it here */
		SeqReader rr(o.pos[0].c_str());
		if(!rr.fp) { fprintf(stderr, "[E::main_align] failed to open index file `%s'. Please check file path and it exists.\n", o.pos[0].c_str()); stop_pinner(); return 1; }
		std::vector<MabIdxSeq> refs; Rec r;
		while(rr.next(r)) { if(r.seq.empty()) { continue; } MabIdxSeq q; q.name = r.name; q.seq = std::move(r.seq); refs.push_back(std::move(q)); }
		std::string err;
		if(!mab_build_index(refs, o.ip, im.own, err)) { fprintf(stderr, "[E::main_index] failed to build index from `%s': %s\n", o.pos[0].c_str(), err.c_str()); stop_pinner(); return 1; }
		fprintf(stderr, "[M::main_index::%.3f] built index for %zu target sequence(s).\n", now() - t0, refs.size());
		im.blob = im.own.data(); im.size = im.own.size();
	}
	if(!o.dump.empty()) {
		if(!mab_write_mai(o.dump.c_str(), std::vector<uint8_t>(im.blob, im.blob + im.size))) { fprintf(stderr, "[E::main_index] failed to write index to `%s'.\n", o.dump.c_str()); stop_pinner(); return 1; }
		fprintf(stderr, "[M::main] Command: %s\n[M::main] Real time: %.3f sec\n", cmdline.c_str(), now() - t0);
		return 0;
	}
	/* Pipeline (the reference's source -> workers -> drain, minialign.c:4565-4643, with the workers' job done by the GPU):
	 *   reader   one thread: the read files in big blocks straight into page-locked chunk buffers, cut at record boundaries
	 *   workers  one thread per context (several contexts per GPU, any number of GPUs): chunk -> mab_text_begin (parse, map) ->
	 *            wait for its turn in file order -> mab_text_commit with the `rlen` word the previous chunk left behind (the
	 *            reference thread's state, DESIGN.md: this keeps the output identical to the -t1 run however the chunks are
	 *            spread) -> mab_text_finish (post-processing, SAM text) into a page-locked output buffer
	 *   writer   one thread: write(2) the SAM text of the chunks in file order
	 * A chunk the device reader does not take (MAB_EFORMAT: wrapped FASTQ, ...) is parsed on the host and goes through the
	 * record-level entry point and the host formatter instead. */
	double t_idx = now() - t0;
	if(o.contexts > 1 && getenv("MAB_EXT_CTAS") == nullptr) { setenv("MAB_EXT_CTAS", "3", 1); }	/* contexts that run side by side launch 3 of the 6 possible extend CTAs per SM each: the other chunks' small kernels need registers to run underneath */
	std::vector<mab_ctx *> ctxs(n_ctx, nullptr);
	{	/* one parent context per device (uploads the index), clones share its image; devices are set up in parallel */
		std::vector<std::thread> th; std::vector<std::string> errs(devices.size());
		std::vector<mab_ctx *> parents(devices.size(), nullptr);
		if(ld) {
			if(mab_load_end(ld, im.blob, im.size, parents.data()) != MAB_OK) { fprintf(stderr, "[E::main_align] failed to instanciate alignment context: %s\n", mab_last_error()); stop_pinner(); return 1; }
			ld = nullptr;
		}
		for(size_t d = 0; d < devices.size(); d++) {
			th.emplace_back([&, d]() {
				mab_ctx *p = parents[d] ? parents[d] : mab_init(im.blob, im.size, &ld_prm, devices[d]);
				if(!p) { errs[d] = mab_last_error(); return; }
				ctxs[d] = p;
				mab_text_reserve(p, chunk_bytes);									/* clones share it: every context sizes its buffers for a full chunk at once */
				for(unsigned c = 1; c < std::max(1u, o.contexts); c++) { mab_ctx *q = mab_clone(p); if(!q) { errs[d] = mab_last_error(); return; } ctxs[c * devices.size() + d] = q; }
				/* what is free next to the index is shared by the contexts of this GPU: 60 % of a context's share for its DP arenas, the
				 * rest for its text, read block, minimizer records, workspaces, result pool and SAM text */
				uint64_t fr = 0, tot = 0;
				if(mab_device_memory(p, &fr, &tot) == MAB_OK && fr > (8ull << 30)) {
					uint64_t share = (fr - (4ull << 30)) / std::max(1u, o.contexts);
					for(unsigned c = 0; c < std::max(1u, o.contexts); c++) { mab_set_arena_budget(ctxs[c * devices.size() + d], share * 6 / 10); }
				}
			});
		}
		for(auto &x : th) { x.join(); }
		for(auto &e : errs) { if(!e.empty()) { fprintf(stderr, "[E::main_align] failed to instanciate alignment context: %s\n", e.c_str()); stop_pinner(); return 1; } }
	}
	{ std::unique_lock<std::mutex> lk(mu); pin_go = true; cv.notify_all(); }
	mab_ctx *ctx0 = ctxs[0];
	uint32_t n_ref = mab_n_ref(ctx0);
	std::vector<MabSamRef> refs(n_ref);
	for(uint32_t i = 0; i < n_ref; i++) { mab_ref_info(ctx0, i, &refs[i].name, &refs[i].l_name, &refs[i].l_seq, &refs[i].seq); }
	fprintf(stderr, "[M::main_align::%.3f] loaded/built index for %u target sequence(s) (index file %.3f s, %u device context(s) on %zu GPU(s) %.3f s).\n", now() - t0, n_ref, t_idx, n_ctx, devices.size(), now() - t0 - t_idx);
	double tmap = now();
	struct stat st_out; const bool out_is_file = fstat(1, &st_out) == 0 && S_ISREG(st_out.st_mode);
	uint64_t out_ofs = out_is_file ? (uint64_t)std::max<off_t>(0, lseek(1, 0, SEEK_CUR)) : 0;	/* regular file: the writers pwrite() concurrently at known offsets */
	uint64_t next_commit = 0, next_write = 0, tot_bases = 0, tot_reads = 0, n_chunks_total = 0;
	uint32_t rlen_chain = 0;												/* a fresh reference thread starts with rlen = 0 (calloc'ed mm_tbuf_t) */
	std::string header;
	{ uint64_t n = mab_sam_header_text(ctx0, "0.6.0-devel", cmdline.c_str(), nullptr, 0); header.resize(n); mab_sam_header_text(ctx0, "0.6.0-devel", cmdline.c_str(), &header[0], n); }
	std::thread reader([&]() {
		uint64_t id = 0;
		std::string carry;													/* the cut-off tail of the previous block: start of the next chunk */
		for(size_t qi = 1; qi < o.pos.size(); qi++) {
			ByteSource src(o.pos[qi].c_str());
			if(!src.ok()) { std::unique_lock<std::mutex> lk(mu); Chunk *c = nullptr; cv.wait(lk, [&]() { return !free_chunks.empty() || failed; }); if(failed) { return; } c = free_chunks.front(); free_chunks.pop_front(); c->len = 0; c->id = id++; c->file = qi; c->open_failed = true; c->last_of_file = true; ready.push_back(c); cv.notify_all(); break; }
			bool eof = false; carry.clear();
			if(src.fd >= 0 && n_ctx > 1) {
				/* a plain file is read by several threads: the chunk boundaries (any record start will do) are found first from a few small
				 * reads around the multiples of the chunk size, then the chunks are pread() side by side -- one thread copies 4-6 GB/s out
				 * of the page cache, which several GPUs outrun.  Buffers are granted and chunks released in id order (a context waits
				 * for its turn in file order, so a later chunk must never hold what an earlier one still needs). */
				struct stat fs; std::vector<uint64_t> cuts;
				char first = 0;
				if(fstat(src.fd, &fs) == 0 && S_ISREG(fs.st_mode) && fs.st_size > 0 && pread(src.fd, &first, 1, 0) == 1 && (first == '>' || first == '@')) {
					const uint64_t F = (uint64_t)fs.st_size, W = 4ull << 20;
					std::vector<char> win(W);
					cuts.push_back(0);
					bool good = true;
					for(uint64_t x = chunk_bytes; x < F && good; x += chunk_bytes) {
						uint64_t lo = std::max(x > W ? x - W : 0, cuts.back());
						int64_t n = pread(src.fd, win.data(), (size_t)std::min<uint64_t>(W, F - lo), (off_t)lo);
						uint64_t c = n > 2 ? record_cut(win.data(), std::min<uint64_t>((uint64_t)n, x - lo), first) : 0;
						if(c == 0 || lo + c <= cuts.back()) { good = false; break; }			/* a record longer than the window, ...: the sequential reader sorts it out */
						cuts.push_back(lo + c);
					}
					uint64_t end = F;															/* blank lines at the end of the file */
					if(good) {
						char tail[64]; int64_t n = pread(src.fd, tail, (size_t)std::min<uint64_t>(64, F), (off_t)(F - std::min<uint64_t>(64, F)));
						while(n > 1 && tail[n - 1] == '\n' && tail[n - 2] == '\n' && end > cuts.back() + 1) { n--; end--; }
						cuts.push_back(end);
					} else { cuts.clear(); }
				}
				if(!cuts.empty()) {
					const size_t nch = cuts.size() - 1;
					size_t next_k = 0; uint64_t next_release = id; std::map<uint64_t, Chunk *> pending; bool rd_failed = false;
					auto rd = [&]() {
						for(;;) {
							Chunk *c; size_t k;
							{
								std::unique_lock<std::mutex> lk(mu);
								cv.wait(lk, [&]() { return !free_chunks.empty() || failed || rd_failed || next_k >= nch; });
								if(failed || rd_failed || next_k >= nch) { return; }
								k = next_k++; c = free_chunks.front(); free_chunks.pop_front();
							}
							const uint64_t len = cuts[k + 1] - cuts[k];
							uint64_t got = 0;
							if(len <= c->cap) { while(got < len) { ssize_t n = pread(src.fd, c->buf + got, (size_t)std::min<uint64_t>(len - got, 1u << 30), (off_t)(cuts[k] + got)); if(n <= 0) { break; } got += (uint64_t)n; } }
							std::unique_lock<std::mutex> lk(mu);
							if(got != len) { rd_failed = true; fprintf(stderr, "[E::main_align] failed to read sequence file `%s'.\n", o.pos[qi].c_str()); failed = true; cv.notify_all(); return; }
							c->len = len; c->id = id + k; c->file = qi; c->last_of_file = k + 1 == nch; c->open_failed = false;
							pending[c->id] = c;
							while(!pending.empty() && pending.begin()->first == next_release) { ready.push_back(pending.begin()->second); pending.erase(pending.begin()); next_release++; }
							cv.notify_all();
						}
					};
					std::vector<std::thread> th; for(unsigned t = 0; t < std::min<unsigned>(4, (unsigned)nch); t++) { th.emplace_back(rd); }
					for(auto &x : th) { x.join(); }
					{ std::unique_lock<std::mutex> lk(mu); if(failed) { return; } }
					id += nch;
					continue;
				}
			}
			while(!eof) {
				Chunk *c;
				{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return !free_chunks.empty() || failed; }); if(failed) { return; } c = free_chunks.front(); free_chunks.pop_front(); }
				uint64_t fill = carry.size();
				if(fill > c->cap) { c->len = 0; }							/* (cannot happen: the carry is cut below the capacity) */
				memcpy(c->buf, carry.data(), fill); carry.clear();
				while(fill < chunk_bytes && !eof) { int64_t n = src.read(c->buf + fill, (chunk_bytes - fill)); if(n <= 0) { eof = true; break; } fill += (uint64_t)n; }
				uint64_t cut = fill;
				if(!eof) {
					cut = record_cut(c->buf, fill);
					while(cut == 0 && !eof) {								/* one record longer than the chunk: keep reading until it ends */
						if(fill + (1 << 20) > c->cap) { break; }
						int64_t n = src.read(c->buf + fill, std::min<uint64_t>(c->cap - fill, 1 << 20)); if(n <= 0) { eof = true; break; } fill += (uint64_t)n;
						cut = eof ? fill : record_cut(c->buf, fill);
					}
					if(eof) { cut = fill; }
					if(cut == 0) { cut = fill; }							/* give up cutting: the device reader will reject what is not whole records */
				}
				carry.assign(c->buf + cut, fill - cut);
				while(cut > 0 && eof && c->buf[cut - 1] == '\n' && cut > 1 && c->buf[cut - 2] == '\n') { cut--; }	/* blank lines at the end of the file */
				c->len = cut; c->id = id++; c->file = qi; c->last_of_file = eof; c->open_failed = false;
				std::unique_lock<std::mutex> lk(mu); ready.push_back(c); cv.notify_all();
			}
		}
		std::unique_lock<std::mutex> lk(mu); read_done = true; n_chunks_total = id; cv.notify_all();
	});
	double t_wr = 0;
	if(write_all(1, header.data(), header.size()) != 0) { fprintf(stderr, "[E::main_align] failed to write the SAM header\n"); stop_pinner(); return 1; }
	out_ofs += header.size();
	auto writer_fn = [&]() {
		while(true) {
			Out *x; uint64_t ofs;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&]() { return done_outs.count(next_write) || failed || (read_done && next_write == n_chunks_total); });
				if(failed || !done_outs.count(next_write)) { return; }
				x = done_outs[next_write]; done_outs.erase(next_write);
				uint64_t len = x->spill.empty() ? x->len : x->spill.size();
				ofs = out_ofs; out_ofs += len;
				if(out_is_file) { next_write++; cv.notify_all(); }			/* the next chunk's writer may start: the offsets are fixed */
			}
			double tm0 = now();
			const char *p = x->spill.empty() ? x->buf : x->spill.data(); uint64_t len = x->spill.empty() ? x->len : x->spill.size();
			int rc = 0;
			if(out_is_file) { while(len) { ssize_t w = pwrite(1, p, len > (1u << 30) ? (1u << 30) : len, (off_t)ofs); if(w < 0) { if(errno == EINTR) { continue; } rc = -1; break; } p += w; len -= (uint64_t)w; ofs += (uint64_t)w; } }
			else { rc = write_all(1, p, len); }
			std::unique_lock<std::mutex> lk(mu);
			t_wr += now() - tm0;
			if(rc != 0) { failed = true; }
			x->spill.clear(); x->len = 0; x->busy = false;
			if(!out_is_file) { next_write++; }
			cv.notify_all();
		}
	};
	std::vector<std::thread> writers;
	for(unsigned i = 0; i < (out_is_file ? std::min(4u, n_ctx) : 1u); i++) { writers.emplace_back(writer_fn); }
	std::vector<double> t_begin(n_ctx, 0), t_turn(n_ctx, 0), t_finish(n_ctx, 0);
	std::atomic<uint64_t> n_redo(0), n_fallback(0), n_failed_reads(0);
	auto worker = [&](unsigned w) {
		mab_ctx *ctx = ctxs[w];
		while(true) {
			Chunk *c; Out *x;
			{
				std::unique_lock<std::mutex> lk(mu);
				auto my_out = [&]() -> Out * { for(unsigned i = 0; i < 2; i++) { if(!out_pool[2 * w + i].busy && out_pool[2 * w + i].buf != nullptr) { return &out_pool[2 * w + i]; } } return nullptr; };
				cv.wait(lk, [&]() { return (!ready.empty() && my_out() != nullptr) || failed || (read_done && ready.empty()); });
				if(failed || ready.empty()) { return; }
				c = ready.front(); ready.pop_front(); x = my_out(); x->busy = true;
			}
			x->id = c->id; x->len = 0; x->spill.clear();
			auto fail = [&](const std::string &msg) { fprintf(stderr, "%s\n", msg.c_str()); std::unique_lock<std::mutex> lk(mu); failed = true; cv.notify_all(); };
			if(c->open_failed) { fail("[E::main_align] failed to open sequence file `" + o.pos[c->file] + "'. Please check file path and format."); return; }
			mab_text_info_t info; memset(&info, 0, sizeof(info));
			double tm0 = now();
			const uint32_t flags = o.tags | (o.keep_qual ? MAB_TEXT_KEEP_QUAL : 0);
			int rc = c->len ? mab_text_begin(ctx, c->buf, c->len, flags, 0, 0, &info) : MAB_OK;
			bool host_path = rc == MAB_EFORMAT;
			if(rc != MAB_OK && !host_path) { fail(std::string("[E::main_align] failed to map sequence file `") + o.pos[c->file] + "': " + mab_last_error()); return; }
			t_begin[w] += now() - tm0; tm0 = now();
			{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return next_commit == c->id || failed; }); if(failed) { return; } }
			t_turn[w] += now() - tm0; tm0 = now();
			uint64_t reads_here = 0, bases_here = 0;
			uint32_t rlen_next = rlen_chain;
			if(host_path) {													/* host reader + record-level mapper + host formatter, in turn (the context's own chain is set by hand) */
				n_fallback++;
				if(!map_chunk_on_host(ctx, c->buf, c->len, o, refs, rlen_chain, &rlen_next, x->spill, &reads_here, &bases_here)) { fail(std::string("[E::main_align] failed to map sequence file `") + o.pos[c->file] + "': " + mab_last_error()); return; }
			} else if(c->len) {
				rc = mab_text_commit(ctx, rlen_chain, &info);
				if(rc != MAB_OK) { fail(std::string("[E::main_align] failed to map sequence file `") + o.pos[c->file] + "': " + mab_last_error()); return; }
				if(info.rlen_valid) { rlen_next = info.rlen_next; }
			}
			{ std::unique_lock<std::mutex> lk(mu); rlen_chain = rlen_next; next_commit++; cv.notify_all(); }
			if(!host_path && c->len) {
				const char *ptr = nullptr;
				if(x->buf == nullptr) { x->cap = c->len + c->len / 2 + (1 << 20); x->buf = (char *)mab_host_alloc(x->cap); if(!x->buf) { fail(std::string("[E::main_align] ") + mab_last_error()); return; } }
				rc = mab_text_finish(ctx, x->buf, x->cap, &ptr, &info);		/* device -> this page-locked buffer; the context is free for the next chunk afterwards */
				if(rc == MAB_ENOMEM && info.sam_bytes > x->cap) {			/* more text than estimated: a larger buffer, format again */
					x->cap = info.sam_bytes + info.sam_bytes / 8 + (1 << 20); x->buf = (char *)mab_host_alloc(x->cap);	/* (the smaller buffer stays behind: it is part of the pinner's arena) */
					if(!x->buf) { fail(std::string("[E::main_align] ") + mab_last_error()); return; }
					rc = mab_text_finish(ctx, x->buf, x->cap, &ptr, &info);
				}
				if(rc != MAB_OK) { fail(std::string("[E::main_align] failed to format the alignments of `") + o.pos[c->file] + "': " + mab_last_error()); return; }
				x->len = info.sam_bytes;
				reads_here = info.n_reads; bases_here = info.n_bases;
				mab_stats_t st; mab_last_stats(ctx, &st); n_redo += st.n_retry; n_failed_reads += st.n_failed;
			}
			t_finish[w] += now() - tm0;
			bool last = c->last_of_file; size_t file = c->file;
			{
				std::unique_lock<std::mutex> lk(mu);
				tot_reads += reads_here; tot_bases += bases_here;
				free_chunks.push_back(c); done_outs[x->id] = x; cv.notify_all();
			}
			if(last) { fprintf(stderr, "[M::main_align::%.3f] finished mapping `%s' onto `%s'.\n", now() - t0, o.pos[file].c_str(), o.pos[0].c_str()); }
		}
	};
	std::vector<std::thread> workers;
	for(unsigned w = 0; w < n_ctx; w++) { workers.emplace_back(worker, w); }
	for(auto &x : workers) { x.join(); }
	{ std::unique_lock<std::mutex> lk(mu); cv.notify_all(); }
	reader.join(); for(auto &x : writers) { x.join(); }
	if(failed) { stop_pinner(); return 1; }
	double sb = 0, st = 0, sf = 0; for(unsigned w = 0; w < n_ctx; w++) { sb += t_begin[w]; st += t_turn[w]; sf += t_finish[w]; }
	fprintf(stderr, "[M::main_align] host pipeline: %u context(s); per context on average: parse+map %.3f s, waiting for its turn %.3f s, post+SAM+copy %.3f s; writer: writing %.3f s; reads re-mapped for the rlen chain: %llu; chunks parsed on the host: %llu\n",
		n_ctx, sb / n_ctx, st / n_ctx, sf / n_ctx, t_wr, (unsigned long long)n_redo.load(), (unsigned long long)n_fallback.load());
	if(n_failed_reads.load()) { fprintf(stderr, "[W::main_align] %llu read(s) overflowed a per-read device structure and are reported unmapped\n", (unsigned long long)n_failed_reads.load()); }
	double tm = now() - tmap;
	fprintf(stderr, "[M::main] mapped %llu reads / %.1f Mbases in %.3f sec (%.1f Mbases/s)\n", (unsigned long long)tot_reads, tot_bases / 1e6, tm, tot_bases / 1e6 / tm);
	fprintf(stderr, "[M::main] Command: %s\n[M::main] Real time: %.3f sec\n", cmdline.c_str(), now() - t0);
	/* everything is written: leave without unpinning the host pools and freeing the DP arenas one by one (0.5-2 s for nothing;
	 * the driver reclaims the context with the process) */
	fflush(stdout); fflush(stderr);
	_exit(0);
}
