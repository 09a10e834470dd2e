// kmc_cli.cpp — `kmer-count`: drop-in for the reference binary (k-mer-count/src/main.rs) over libkmc's C ABI.
//
//   kmer-count                      exactly the reference: read ./sample.fasta (main.rs:44), count the L27‖R27
//                                   gapped chunks (main.rs:48-49,63-80), print them sorted, duplicates repeated,
//                                   one per line (main.rs:87-90); exit 101 with a panic-style message on the
//                                   inputs on which the reference panics (main.rs:44,59,23,35).
//   kmer-count FASTA -k K [-o OUT]  ordinary canonical k-mers, "kmer<TAB>count" lines ascending by k-mer.
//   kmer-count --gpus N ...         the same job on N GPUs of this box, the same bytes out: hands over to the
//                                   multi-process host program (k-mer-count_b200/cli_dist.py under torchrun, one
//                                   process per GPU; records sharded, keys routed to their owner GPU, one merged stream).
//
// Host side only: FASTA parsing (what bio::io::fasta::Reader does for main.rs:45,59-62), pinned staging,
// text output.  All counting happens on the GPU behind kmc.h; there is no CPU counting path in this program.
#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#include "../../include/kmc.h"

namespace {

struct Options {
  std::string fasta = "sample.fasta"; // main.rs:44
  std::string out;                    // empty = stdout (main.rs:88-90)
  std::string stats;
  uint32_t mode = KMC_MODE_LR_GAPPED; // no arguments = the reference's computation
  uint32_t k = 31;
  uint32_t canonical = 0;
  uint32_t strategy = KMC_STRATEGY_AUTO;
  uint32_t l_len = 0, r_len = 0, d_min = 0, d_max = 0;
  int device = -1;
  int parts = 1;           // --parts P: count in P key-range passes (kmc_finish_part)
  bool host_parse = false; // --host-parse: parse FASTA on the host (default: on the device, kmc_submit_fasta)
  bool expanded = true; // lr-gapped: repeat each key `count` times (the reference's output); --counts switches it off
};

[[noreturn]] void panic(const char *what, const std::string &detail) {
  // the reference's failures are Rust panics: message on stderr, exit status 101
  fprintf(stderr, "thread 'main' panicked: %s%s%s\n", what, detail.empty() ? "" : ": ", detail.c_str());
  exit(101);
}

void usage() {
  fputs("usage: kmer-count [FASTA] [-k K] [-o OUT] [--mode lr-gapped|contiguous] [--canonical|--no-canonical]\n"
        "                  [--strategy auto|hash|sort|baseline] [--lr L R DMIN DMAX] [--counts] [--device N] [--stats FILE]\n"
        "                  [--host-parse] [--parts P] [--gpus N]\n"
        "  no arguments: read ./sample.fasta and print the reference's output (sorted L27+R27 gapped chunks)\n",
        stderr);
}

// --gpus N (N > 1): one process per GPU.  This program is one process on one GPU; the multi-GPU job is run by the
// Python host side over the same C ABI (cli_dist.py), launched with torchrun as a child.  Does not return when it hands over.
void maybe_hand_over_to_ranks(int argc, char **argv) {
  int gpus = 1, at = -1;
  for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], "--gpus")) { gpus = atoi(argv[i + 1]); at = i; }
  if (at < 0) return;
  if (gpus < 1) { fputs("kmer-count: --gpus needs a positive number\n", stderr); exit(2); }
  if (gpus == 1) return;
  char exe[4096];
  ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 1);
  if (n <= 0) { perror("kmer-count: readlink /proc/self/exe"); exit(3); }
  exe[n] = 0;
  std::string root(exe); // <root>/k-mer-count_b200/bin/kmer-count
  for (int up = 0; up < 3; up++) { size_t p = root.rfind('/'); if (p == std::string::npos) break; root.resize(p); }
  std::string pp = root;
  if (const char *old = getenv("PYTHONPATH")) { pp += ":"; pp += old; }
  setenv("PYTHONPATH", pp.c_str(), 1);
  const std::string nproc = std::to_string(gpus), port = std::to_string(29600 + (int)(getpid() % 300));
  std::vector<std::string> a = {"python3", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", nproc, "--master-addr", "127.0.0.1",
                                "--master-port", port, "-m", "kmer_count_b200.cli_dist"};
  for (int i = 1; i < argc; i++) { if (i == at || i == at + 1) continue; a.push_back(argv[i]); }
  std::vector<char *> av;
  for (auto &x : a) av.push_back(const_cast<char *>(x.c_str()));
  av.push_back(nullptr);
  // torchrun turns any failing rank into its own exit status 1, so the ranks report the program's status — 101 when
  // the reference would have panicked — through a file and leave with 0
  char status_path[] = "/tmp/kmer-count-status-XXXXXX";
  int sfd = mkstemp(status_path);
  if (sfd >= 0) { close(sfd); setenv("KMC_CLI_STATUS", status_path, 1); }
  pid_t pid = fork();
  if (pid == 0) {
    execvp(av[0], av.data());
    perror("kmer-count: cannot start python3 -m torch.distributed.run");
    _exit(3);
  }
  int st = 0, code = 3;
  if (pid > 0 && waitpid(pid, &st, 0) == pid) code = WIFEXITED(st) ? WEXITSTATUS(st) : 3;
  if (sfd >= 0) {
    if (FILE *f = fopen(status_path, "r")) { int v = -1; if (fscanf(f, "%d", &v) == 1 && v >= 0) code = v; fclose(f); }
    unlink(status_path);
  }
  exit(code);
}

Options parse_args(int argc, char **argv) {
  Options o;
  bool mode_given = false, canon_given = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto need = [&](int n) { if (i + n >= argc) { usage(); exit(2); } };
    if (a == "-k") { need(1); o.k = (uint32_t)atoi(argv[++i]); if (!mode_given) o.mode = KMC_MODE_CONTIGUOUS; }
    else if (a == "-o") { need(1); o.out = argv[++i]; }
    else if (a == "--stats") { need(1); o.stats = argv[++i]; }
    else if (a == "--device") { need(1); o.device = atoi(argv[++i]); }
    else if (a == "--mode") {
      need(1); std::string m = argv[++i]; mode_given = true;
      if (m == "lr-gapped") o.mode = KMC_MODE_LR_GAPPED; else if (m == "contiguous") o.mode = KMC_MODE_CONTIGUOUS; else { usage(); exit(2); }
    }
    else if (a == "--canonical") { o.canonical = 1; canon_given = true; }
    else if (a == "--no-canonical") { o.canonical = 0; canon_given = true; }
    else if (a == "--counts") o.expanded = false;
    else if (a == "--host-parse") o.host_parse = true;
    else if (a == "--gpus") { need(1); ++i; } // 1: this process (larger values never get here)
    else if (a == "--parts") { need(1); o.parts = atoi(argv[++i]); if (o.parts < 1) { usage(); exit(2); } }
    else if (a == "--strategy") {
      need(1); std::string s = argv[++i];
      o.strategy = s == "hash" ? KMC_STRATEGY_HASH : s == "sort" ? KMC_STRATEGY_SORT : s == "baseline" ? KMC_STRATEGY_SORT_BASELINE : KMC_STRATEGY_AUTO;
    }
    else if (a == "--lr") { need(4); o.l_len = atoi(argv[++i]); o.r_len = atoi(argv[++i]); o.d_min = atoi(argv[++i]); o.d_max = atoi(argv[++i]); }
    else if (a == "-h" || a == "--help") { usage(); exit(0); }
    else if (!a.empty() && a[0] != '-') o.fasta = a;
    else { usage(); exit(2); }
  }
  if (o.mode == KMC_MODE_CONTIGUOUS && !canon_given) o.canonical = 1;
  if (o.mode == KMC_MODE_LR_GAPPED) o.canonical = 0;
  return o;
}

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// bio 0.41 fasta::Reader as main.rs:58-62 drives it: header line starts with '>', sequence lines are appended with
// trailing whitespace trimmed, a first line without '>' is an error, and the loop ends at the first record whose id,
// description and sequence are all empty.  Fills the staging buffers batch by batch and submits them.
struct Feeder {
  kmc_ctx *ctx;
  uint8_t *bases = nullptr;
  uint64_t *off = nullptr;
  size_t cap_b = 0, cap_r = 0, nb = 0, nr = 0;
  uint64_t total_bases = 0, total_recs = 0;
  void acquire(size_t want_b, size_t want_r) {
    int rc = kmc_staging(ctx, want_b, want_r, &bases, &off, &cap_b, &cap_r);
    if (rc) panic("kmc_staging", kmc_last_error(ctx));
    nb = nr = 0;
    off[0] = 0;
  }
  void flush() {
    if (!nr) return;
    int rc = kmc_submit(ctx, nb, nr);
    if (rc) panic("kmc_submit", kmc_last_error(ctx));
    total_bases += nb; total_recs += nr;
    acquire(cap_b, cap_r);
  }
};

void feed_fasta(const Options &o, kmc_ctx *ctx, Feeder &fd) {
  FILE *f = fopen(o.fasta.c_str(), "rb");
  if (!f) panic("Error during opening the file", strerror(errno)); // main.rs:44
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> buf((size_t)(sz > 0 ? sz : 0));
  if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) panic("read failed", o.fasta);
  fclose(f);
  if (!o.host_parse) { // the file's bytes go to the GPU as they are; records are found there
    int rc = kmc_submit_fasta(ctx, buf.data(), buf.size(), &fd.total_bases, &fd.total_recs);
    if (rc == KMC_E_FORMAT) panic("called `Result::unwrap()` on an `Err` value", "Expected > at record start."); // main.rs:59
    if (rc) panic("kmc_submit_fasta", kmc_last_error(ctx));
    return;
  }
  const size_t batch = (size_t)256 << 20;
  fd.ctx = ctx;
  fd.acquire(std::min<size_t>(batch, (size_t)sz + 1024), 1 << 16);
  long p = 0;
  while (p < sz) {
    long e = p;
    while (e < sz && buf[e] != '\n') e++;
    if (buf[p] != '>') panic("called `Result::unwrap()` on an `Err` value", "Expected > at record start."); // main.rs:59
    long he = e;
    while (he > p + 1 && is_space(buf[he - 1])) he--;
    const bool header_empty = he == p + 1;
    p = e < sz ? e + 1 : sz;
    // measure the record first so that it never straddles two batches
    long q = p;
    size_t len = 0;
    while (q < sz && buf[q] != '>') {
      long le = q;
      while (le < sz && buf[le] != '\n') le++;
      long te = le;
      while (te > q && is_space(buf[te - 1])) te--;
      len += (size_t)(te - q);
      q = le < sz ? le + 1 : sz;
    }
    if (header_empty && len == 0) break; // record.is_empty() → main.rs:60-62
    if (fd.nb + len > fd.cap_b || fd.nr + 1 > fd.cap_r) {
      fd.flush();
      if (len > fd.cap_b) fd.acquire(len, fd.cap_r);
    }
    q = p;
    while (q < sz && buf[q] != '>') {
      long le = q;
      while (le < sz && buf[le] != '\n') le++;
      long te = le;
      while (te > q && is_space(buf[te - 1])) te--;
      memcpy(fd.bases + fd.nb, buf.data() + q, (size_t)(te - q));
      fd.nb += (size_t)(te - q);
      q = le < sz ? le + 1 : sz;
    }
    fd.off[++fd.nr] = fd.nb;
    p = q;
  }
  fd.flush();
}

// table rows → text, formatted on the device (kmc_format) chunk by chunk.  expanded: each key `count` times, one per
// line (main.rs:88-90); else "kmer\tcount".
void emit(const Options &o, kmc_ctx *ctx, uint64_t n_distinct, FILE *out) {
  const int expanded = o.mode == KMC_MODE_LR_GAPPED && o.expanded;
  const size_t max_bytes = (size_t)1 << 30;
  uint64_t chunk = 1 << 22;
  for (uint64_t first = 0; first < n_distinct;) {
    uint64_t n = std::min<uint64_t>(chunk, n_distinct - first);
    const char *text = nullptr;
    size_t len = 0;
    int rc = kmc_format(ctx, first, n, expanded, max_bytes, &text, &len);
    if (rc == KMC_E_CAPACITY && n > 1) { chunk = std::max<uint64_t>(1, n / 4); continue; } // huge multiplicities: fewer rows at a time
    if (rc == KMC_E_CAPACITY) { // a single row whose expansion exceeds the buffer: expand it here
      uint64_t lo = 0, hi = 0, cnt = 0;
      if (kmc_read(ctx, first, 1, &lo, &hi, &cnt)) panic("kmc_read", kmc_last_error(ctx));
      const uint32_t nb = kmc_key_bases(ctx);
      char line[160];
      unsigned __int128 v = ((unsigned __int128)hi << 64) | lo;
      for (uint32_t b = 0; b < nb; b++) { line[nb - 1 - b] = "ACGT"[(unsigned)(v & 3)]; v >>= 2; }
      line[nb] = '\n';
      for (uint64_t c = 0; c < cnt; c++) fwrite(line, 1, nb + 1, out);
      first += 1;
      continue;
    }
    if (rc) panic("kmc_format", kmc_last_error(ctx));
    fwrite(text, 1, len, out);
    first += n;
  }
}

} // namespace

int main(int argc, char **argv) {
  maybe_hand_over_to_ranks(argc, argv);
  Options o = parse_args(argc, argv);
  kmc_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = KMC_ABI_VERSION;
  cfg.mode = o.mode; cfg.k = o.k; cfg.canonical = o.canonical; cfg.strategy = o.strategy; cfg.device = o.device;
  cfg.l_len = o.l_len; cfg.r_len = o.r_len; cfg.d_min = o.d_min; cfg.d_max = o.d_max;
  kmc_ctx *ctx = nullptr;
  int rc = kmc_create(&ctx, &cfg);
  if (rc) { fprintf(stderr, "kmer-count: %s (%s)\n", kmc_last_error(nullptr), kmc_strerror(rc)); return 3; }
  Feeder fd;
  feed_fasta(o, ctx, fd);
  FILE *out = nullptr;
  // --parts P: the key space in P ascending ranges, one counting pass each (inputs whose keys exceed HBM); the output
  // is the same text, since the ranges' tables follow each other in key order
  for (uint32_t part = 0; part < (uint32_t)o.parts; part++) {
    uint64_t n_distinct = 0, n_total = 0;
    rc = kmc_finish_part(ctx, part, (uint32_t)o.parts, &n_distinct, &n_total);
    if (rc == KMC_E_BADBASE) panic("Unexpected charactor appears in a chunk", kmc_last_error(ctx));                  // main.rs:23
    if (rc == KMC_E_EMPTY) panic("index out of bounds: the len is 0 but the index is 0", kmc_last_error(ctx));        // main.rs:35
    if (rc) panic(kmc_strerror(rc), kmc_last_error(ctx));
    if (!out) { // opened after the first count: the reference prints nothing when it panics
      out = o.out.empty() ? stdout : fopen(o.out.c_str(), "wb");
      if (!out) panic("cannot open output", o.out);
    }
    emit(o, ctx, n_distinct, out);
  }
  if (out == stdout) fflush(out); else if (out) fclose(out);
  if (!o.stats.empty()) {
    size_t n = kmc_stats_json(ctx, nullptr, 0);
    std::vector<char> s(n);
    kmc_stats_json(ctx, s.data(), n);
    FILE *sf = fopen(o.stats.c_str(), "w");
    if (sf) { fprintf(sf, "%s\n", s.data()); fclose(sf); }
  }
  kmc_destroy(ctx);
  return 0;
}
