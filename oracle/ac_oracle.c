/*
 * ac_oracle.c -- ORACLE: TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU-only restatement of the algorithm the reference runs for its
 * hot path.  The reference (/root/reference/src/lib.rs) holds no matching
 * arithmetic: it calls the third-party crate `aho-corasick` 1.1.4
 * (Cargo.lock:5-12), which is NOT vendored under /root/reference.  This file
 * restates that crate's published algorithm (noncontiguous NFA: trie +
 * BFS failure links with the leftmost "dead state" rule, then a dense DFA
 * tabulated from it, then the find / find_overlapping loops), anchored on the
 * reference's call sites:
 *
 *   build            src/lib.rs:186-215 (str), 401-406 (bytes)
 *   iterator choice  src/lib.rs:42-68   (get_matches: overlapping vs not,
 *                                        MatchError -> ValueError before any byte)
 *   non-overlapping  src/lib.rs:58-60   (try_find_iter, drained at 238-248/261/433)
 *   overlapping      src/lib.rs:52-54   (try_find_overlapping_iter)
 *   tuple layout     src/lib.rs:240-246 (str: code points), 431 (bytes: byte offsets)
 *   byte->code point src/lib.rs:73-88
 *
 * PARITY STATUS: partially pinned.  The Rust reference cannot be built or
 * imported in this image (no rustc/cargo/maturin, no wheel), so this oracle is
 * pinned only against the golden vectors the reference's README and tests
 * hold (tests/golden/reference_vectors.json) and against a brute-force
 * statement of the match semantics (tests/spec_bruteforce.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (ahocorasick_rs_b200)
 * never does.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_DEAD 0u  /* every transition leads back to DEAD */
#define ORC_FAIL 1u  /* sentinel "no edge here, follow the failure link" */
#define ORC_START 2u /* unanchored start state */

enum { ORC_STANDARD = 0, ORC_LEFTMOST_FIRST = 1, ORC_LEFTMOST_LONGEST = 2 };

typedef struct {
    uint32_t next, link;
    uint8_t byte;
} orc_edge;

typedef struct {
    uint32_t pid, link;
} orc_mnode;

typedef struct orc_ac {
    int kind;
    uint32_t nstates, cap_states;
    uint32_t *first_edge; /* head of this state's edge list (sorted by byte); 0 = none */
    uint32_t *fail;
    uint32_t *first_match, *last_match; /* match list head/tail; 0 = none */
    orc_edge *edges;
    uint32_t nedges, cap_edges;
    orc_mnode *mlist;
    uint32_t nm, cap_m;
    uint32_t npatterns;
    uint32_t *pat_len;
    uint32_t max_len;
    /* dense DFA tabulated from the NFA */
    uint8_t classes[256];
    uint32_t ncls, stride;
    uint32_t *trans; /* nstates * stride, entries are state ids */
    /* the same table the way the crate's DFA stores it for speed: ids premultiplied
     * by the stride, match states recognisable without a second lookup (here: top bit) */
    uint32_t *ftrans;
} orc_ac;

static void *xrealloc(void *p, size_t n) {
    void *q = realloc(p, n);
    if (!q) {
        fprintf(stderr, "oracle: out of memory\n");
        abort();
    }
    return q;
}

static uint32_t alloc_state(orc_ac *a) {
    if (a->nstates == a->cap_states) {
        a->cap_states = a->cap_states ? a->cap_states * 2 : 1024;
        a->first_edge = xrealloc(a->first_edge, a->cap_states * sizeof(uint32_t));
        a->fail = xrealloc(a->fail, a->cap_states * sizeof(uint32_t));
        a->first_match = xrealloc(a->first_match, a->cap_states * sizeof(uint32_t));
        a->last_match = xrealloc(a->last_match, a->cap_states * sizeof(uint32_t));
    }
    uint32_t s = a->nstates++;
    a->first_edge[s] = 0;
    a->fail[s] = ORC_START;
    a->first_match[s] = a->last_match[s] = 0;
    return s;
}

/* One trie step with no failure handling.  DEAD absorbs; the start state
 * loops to itself on bytes without an edge (the unanchored-search loop). */
static uint32_t nfa_goto(const orc_ac *a, uint32_t s, uint8_t b) {
    if (s == ORC_DEAD) return ORC_DEAD;
    for (uint32_t e = a->first_edge[s]; e; e = a->edges[e].link) {
        if (a->edges[e].byte == b) return a->edges[e].next;
        if (a->edges[e].byte > b) break;
    }
    return s == ORC_START ? ORC_START : ORC_FAIL;
}

static void add_edge(orc_ac *a, uint32_t s, uint8_t b, uint32_t next) {
    if (a->nedges == a->cap_edges) {
        a->cap_edges = a->cap_edges ? a->cap_edges * 2 : 4096;
        a->edges = xrealloc(a->edges, a->cap_edges * sizeof(orc_edge));
    }
    uint32_t id = a->nedges++;
    a->edges[id].byte = b;
    a->edges[id].next = next;
    uint32_t *slot = &a->first_edge[s];
    while (*slot && a->edges[*slot].byte < b) slot = &a->edges[*slot].link;
    a->edges[id].link = *slot;
    *slot = id;
}

static void add_match(orc_ac *a, uint32_t s, uint32_t pid) {
    if (a->nm == a->cap_m) {
        a->cap_m = a->cap_m ? a->cap_m * 2 : 4096;
        a->mlist = xrealloc(a->mlist, a->cap_m * sizeof(orc_mnode));
    }
    uint32_t id = a->nm++;
    a->mlist[id].pid = pid;
    a->mlist[id].link = 0;
    if (a->last_match[s])
        a->mlist[a->last_match[s]].link = id;
    else
        a->first_match[s] = id;
    a->last_match[s] = id;
}

/* Append src's match list to dst's (own patterns stay first). */
static void copy_matches(orc_ac *a, uint32_t src, uint32_t dst) {
    for (uint32_t m = a->first_match[src]; m; m = a->mlist[m].link) add_match(a, dst, a->mlist[m].pid);
}

static uint32_t nfa_next(const orc_ac *a, uint32_t s, uint8_t b) {
    for (;;) {
        uint32_t n = nfa_goto(a, s, b);
        if (n != ORC_FAIL) return n;
        s = a->fail[s];
    }
}

void orc_free(orc_ac *a) {
    if (!a) return;
    free(a->first_edge);
    free(a->fail);
    free(a->first_match);
    free(a->last_match);
    free(a->edges);
    free(a->mlist);
    free(a->pat_len);
    free(a->trans);
    free(a->ftrans);
    free(a);
}

/*
 * Build.  patterns = blob[offs[i] .. offs[i+1]) for i in [0, n).  Pattern ids
 * are input order (src/lib.rs:186-215 feeds them to the builder in iteration
 * order).  Returns NULL and writes err on failure; empty patterns are rejected
 * by the reference before the crate sees them (src/lib.rs:204-207, 386-389).
 */
orc_ac *orc_build(const uint8_t *blob, const uint64_t *offs, uint64_t n, int kind, char *err, size_t errlen) {
    orc_ac *a = calloc(1, sizeof(orc_ac));
    a->kind = kind;
    a->npatterns = (uint32_t)n;
    a->pat_len = calloc(n ? n : 1, sizeof(uint32_t));
    /* edge / match-node index 0 is the "none" sentinel */
    a->cap_edges = 4096;
    a->edges = xrealloc(NULL, a->cap_edges * sizeof(orc_edge));
    a->nedges = 1;
    a->cap_m = 4096;
    a->mlist = xrealloc(NULL, a->cap_m * sizeof(orc_mnode));
    a->nm = 1;
    alloc_state(a); /* DEAD */
    alloc_state(a); /* FAIL */
    alloc_state(a); /* START */
    a->fail[ORC_START] = ORC_START;

    /* 1. trie */
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *p = blob + offs[i];
        uint64_t len = offs[i + 1] - offs[i];
        if (len == 0) {
            snprintf(err, errlen, "empty pattern at index %llu", (unsigned long long)i);
            orc_free(a);
            return NULL;
        }
        a->pat_len[i] = (uint32_t)len;
        if (len > a->max_len) a->max_len = (uint32_t)len;
        uint32_t prev = ORC_START;
        int saw_match = 0, dropped = 0;
        for (uint64_t d = 0; d < len; d++) {
            /* leftmost-first: an earlier pattern that is a proper prefix of
             * this one always wins, so this one is never inserted */
            saw_match = saw_match || a->first_match[prev] != 0;
            if (kind == ORC_LEFTMOST_FIRST && saw_match) {
                dropped = 1;
                break;
            }
            uint32_t nx = nfa_goto(a, prev, p[d]);
            if (nx == ORC_FAIL || (prev == ORC_START && nx == ORC_START)) {
                nx = alloc_state(a);
                add_edge(a, prev, p[d], nx);
            }
            prev = nx;
        }
        if (!dropped) add_match(a, prev, (uint32_t)i);
    }

    /* 2. failure links, breadth first */
    int leftmost = kind != ORC_STANDARD;
    uint32_t *queue = malloc(sizeof(uint32_t) * a->nstates);
    uint32_t qh = 0, qt = 0;
    for (uint32_t e = a->first_edge[ORC_START]; e; e = a->edges[e].link) {
        uint32_t c = a->edges[e].next;
        queue[qt++] = c;
        a->fail[c] = ORC_START;
        /* a match state right after the start state must never fall back to
         * the start state under leftmost semantics */
        if (leftmost && a->first_match[c]) a->fail[c] = ORC_DEAD;
    }
    while (qh < qt) {
        uint32_t id = queue[qh++];
        for (uint32_t e = a->first_edge[id]; e; e = a->edges[e].link) {
            uint32_t c = a->edges[e].next;
            uint8_t b = a->edges[e].byte;
            queue[qt++] = c;
            if (leftmost && a->first_match[c]) {
                /* trie-terminal state: nothing that starts later may be
                 * found through it or anything below it */
                a->fail[c] = ORC_DEAD;
                continue;
            }
            uint32_t f = a->fail[id];
            while (nfa_goto(a, f, b) == ORC_FAIL) f = a->fail[f];
            f = nfa_goto(a, f, b);
            a->fail[c] = f;
            copy_matches(a, f, c);
        }
    }
    free(queue);

    /* 3. dense DFA over byte classes (bytes no pattern uses share class 0) */
    int used[256] = {0};
    for (uint64_t i = 0; i < offs[n]; i++) used[blob[i]] = 1;
    a->ncls = 1;
    for (int b = 0; b < 256; b++) a->classes[b] = used[b] ? (uint8_t)(a->ncls++) : 0;
    if (a->ncls > 255) { /* all 256 bytes used: class ids must still fit u8 */
        a->ncls = 256;
        for (int b = 0; b < 256; b++) a->classes[b] = (uint8_t)b;
    }
    a->stride = 1;
    while (a->stride < a->ncls) a->stride <<= 1;
    uint64_t cells = (uint64_t)a->nstates * a->stride;
    if (cells > (1ull << 31)) {
        snprintf(err, errlen, "oracle DFA too large (%llu cells)", (unsigned long long)cells);
        orc_free(a);
        return NULL;
    }
    a->trans = calloc(cells, sizeof(uint32_t));
    uint8_t rep[256];
    for (int b = 255; b >= 0; b--) rep[a->classes[b]] = (uint8_t)b;
    for (uint32_t s = 0; s < a->nstates; s++) {
        if (s == ORC_FAIL) continue;
        for (uint32_t c = 0; c < a->ncls; c++) a->trans[(uint64_t)s * a->stride + c] = nfa_next(a, s, rep[c]);
    }
    a->ftrans = calloc(cells, sizeof(uint32_t));
    for (uint64_t i = 0; i < cells; i++) {
        uint32_t t = a->trans[i];
        a->ftrans[i] = (t * a->stride) | (a->first_match[t] ? 0x80000000u : 0u);
    }
    return a;
}

uint32_t orc_num_states(const orc_ac *a) { return a->nstates; }
uint32_t orc_max_pattern_len(const orc_ac *a) { return a->max_len; }

static inline uint32_t step(const orc_ac *a, int use_dfa, uint32_t s, uint8_t b) {
    return use_dfa ? a->trans[(uint64_t)s * a->stride + a->classes[b]] : nfa_next(a, s, b);
}

/*
 * Drain of the reference's iterator (src/lib.rs:42-68 + 238-248 / 433).
 * Writes up to cap (pid, start, end) triples in iteration order; returns the
 * TOTAL number of matches (may exceed cap), or -1 when overlapping is asked of
 * a non-Standard automaton (the crate refuses at iterator creation, before any
 * byte is read: src/lib.rs:52-54 -> match_error_to_pyerror 36-39).
 * use_dfa selects the tabulated DFA or the NFA walk; both must agree.
 */
int64_t orc_find_iter(const orc_ac *a, const uint8_t *hay, uint64_t len, int overlapping, int use_dfa,
                      uint32_t *out_pid, uint64_t *out_start, uint64_t *out_end, uint64_t cap) {
    uint64_t n = 0;
    if (overlapping && a->kind != ORC_STANDARD) return -1;
    if (use_dfa && !overlapping) {
        /* same loop as below on the premultiplied, flagged table (what the timed baseline runs) */
        const uint32_t *T = a->ftrans;
        const uint8_t *cls = a->classes;
        const uint32_t root = ORC_START * a->stride;
        uint64_t start = 0;
        while (start <= len) {
            uint32_t sid = root;
            int have = 0;
            uint32_t mpid = 0;
            uint64_t mend = 0;
            for (uint64_t at = start; at < len; at++) {
                uint32_t e = T[sid + cls[hay[at]]];
                sid = e & 0x7fffffffu;
                if (e >> 31) {
                    have = 1;
                    mpid = a->mlist[a->first_match[sid / a->stride]].pid;
                    mend = at + 1;
                    if (a->kind == ORC_STANDARD) break;
                } else if (sid == ORC_DEAD) {
                    break;
                }
            }
            if (!have) break;
            if (n < cap) {
                out_pid[n] = mpid;
                out_start[n] = mend - a->pat_len[mpid];
                out_end[n] = mend;
            }
            n++;
            start = mend;
        }
        return (int64_t)n;
    }
    if (overlapping) {
        uint32_t sid = ORC_START;
        for (uint64_t at = 0; at < len; at++) {
            sid = step(a, use_dfa, sid, hay[at]);
            /* every pattern on the state's list, in list order, then move on */
            for (uint32_t m = a->first_match[sid]; m; m = a->mlist[m].link) {
                uint32_t pid = a->mlist[m].pid;
                if (n < cap) {
                    out_pid[n] = pid;
                    out_start[n] = at + 1 - a->pat_len[pid];
                    out_end[n] = at + 1;
                }
                n++;
            }
        }
        return (int64_t)n;
    }
    uint64_t start = 0;
    while (start <= len) {
        /* one try_find from `start`, beginning in the start state */
        uint32_t sid = ORC_START;
        int have = 0;
        uint32_t mpid = 0;
        uint64_t mend = 0;
        for (uint64_t at = start; at < len; at++) {
            sid = step(a, use_dfa, sid, hay[at]);
            if (sid == ORC_DEAD) break;
            if (a->first_match[sid]) {
                have = 1;
                mpid = a->mlist[a->first_match[sid]].pid;
                mend = at + 1;
                if (a->kind == ORC_STANDARD) break; /* earliest match wins */
            }
        }
        if (!have) break;
        if (n < cap) {
            out_pid[n] = mpid;
            out_start[n] = mend - a->pat_len[mpid];
            out_end[n] = mend;
        }
        n++;
        start = mend; /* patterns are non-empty, so this always advances */
    }
    return (int64_t)n;
}

/*
 * src/lib.rs:73-88: byte offset -> code point index; entries that are not a
 * char boundary hold UINT64_MAX; entry [len] = number of code points (only
 * written for non-empty input, like the reference).  out has len+1 slots.
 */
void orc_byte_to_code_point(const uint8_t *hay, uint64_t len, uint64_t *out) {
    for (uint64_t i = 0; i <= len; i++) out[i] = UINT64_MAX;
    uint64_t cp = 0;
    for (uint64_t i = 0; i < len; i++)
        if ((hay[i] & 0xC0) != 0x80) out[i] = cp++;
    if (len) out[len] = cp;
}

/* ---- batch driver, used for parity on batches and as the timed CPU baseline ---- */

typedef struct {
    const orc_ac *a;
    const uint8_t *bytes;
    const int64_t *offs;
    int64_t lo, hi;
    int overlapping, codepoints;
    /* outputs (optional): per-haystack counts and a private record buffer */
    uint32_t *counts;
    uint32_t *rec; /* 4 x u32 per match: hay, pid, start, end */
    uint64_t rec_cap, nrec;
    uint64_t total;
    int reps; /* timing only: scan the shard this many times inside one thread launch */
} orc_job;

static void *batch_worker(void *arg) {
    orc_job *j = arg;
    const orc_ac *a = j->a;
    uint32_t pid[64];
    uint64_t st[64], en[64];
    uint64_t *b2c = NULL;
    uint64_t b2c_cap = 0;
    for (int rep = 0; rep < j->reps; rep++)
    for (int64_t h = j->lo; h < j->hi; h++) {
        const uint8_t *hay = j->bytes + j->offs[h];
        uint64_t len = (uint64_t)(j->offs[h + 1] - j->offs[h]);
        if (j->codepoints) { /* the reference builds the map for every call: src/lib.rs:235 */
            if (len + 1 > b2c_cap) {
                b2c_cap = (len + 1) * 2;
                b2c = xrealloc(b2c, b2c_cap * sizeof(uint64_t));
            }
            orc_byte_to_code_point(hay, len, b2c);
        }
        uint64_t cap = 64;
        uint32_t *ppid = pid;
        uint64_t *pst = st, *pen = en;
        int64_t n = orc_find_iter(a, hay, len, j->overlapping, 1, ppid, pst, pen, cap);
        if (n > (int64_t)cap) { /* rare: redo with room for everything */
            cap = (uint64_t)n;
            ppid = malloc(cap * sizeof(uint32_t));
            pst = malloc(cap * sizeof(uint64_t));
            pen = malloc(cap * sizeof(uint64_t));
            orc_find_iter(a, hay, len, j->overlapping, 1, ppid, pst, pen, cap);
        }
        if (n < 0) n = 0;
        if (j->counts) j->counts[h] = (uint32_t)n;
        for (int64_t k = 0; k < n; k++) {
            uint64_t s = pst[k], e = pen[k];
            if (j->codepoints) {
                s = b2c[s];
                e = b2c[e];
            }
            if (j->rec && j->nrec < j->rec_cap) {
                uint32_t *r = j->rec + 4 * j->nrec;
                r[0] = (uint32_t)h;
                r[1] = ppid[k];
                r[2] = (uint32_t)s;
                r[3] = (uint32_t)e;
            }
            j->nrec++;
        }
        j->total += (uint64_t)n;
        if (ppid != pid) {
            free(ppid);
            free(pst);
            free(pen);
        }
    }
    free(b2c);
    return NULL;
}

/*
 * Scan haystacks [0, n) = bytes[offs[h] .. offs[h+1]) with `nthreads` host
 * threads, each owning a contiguous range of haystacks (legitimate because the
 * reference releases the GIL around the scan, src/lib.rs:238/433, and the
 * automaton is immutable).  counts (n entries) and rec (4*rec_cap u32) may be
 * NULL.  Records come back in haystack order, then iteration order.  Returns
 * the total number of matches.
 */
uint64_t orc_scan_batch_reps(const orc_ac *a, const uint8_t *bytes, const int64_t *offs, int64_t n, int overlapping,
                             int codepoints, int nthreads, uint32_t *counts, uint32_t *rec, uint64_t rec_cap, int reps);

uint64_t orc_scan_batch(const orc_ac *a, const uint8_t *bytes, const int64_t *offs, int64_t n, int overlapping,
                        int codepoints, int nthreads, uint32_t *counts, uint32_t *rec, uint64_t rec_cap) {
    return orc_scan_batch_reps(a, bytes, offs, n, overlapping, codepoints, nthreads, counts, rec, rec_cap, 1);
}

/* Same, with every thread scanning its shard `reps` times (the CPU-baseline
 * timing loop: keeps thread start-up out of the measurement).  The returned
 * total counts every repetition. */
uint64_t orc_scan_batch_reps(const orc_ac *a, const uint8_t *bytes, const int64_t *offs, int64_t n, int overlapping,
                             int codepoints, int nthreads, uint32_t *counts, uint32_t *rec, uint64_t rec_cap, int reps) {
    if (reps < 1) reps = 1;
    if (rec) reps = 1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (rec) nthreads = 1; /* ordered records: keep it simple, single writer */
    orc_job jobs[256];
    pthread_t th[256];
    /* balance by bytes */
    int64_t total_bytes = offs[n] - offs[0];
    int64_t h = 0;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (orc_job){a, bytes, offs, h, h, overlapping, codepoints, counts, rec, rec_cap, 0, 0, reps};
        int64_t target = offs[0] + (total_bytes * (t + 1)) / nthreads;
        while (h < n && (offs[h + 1] <= target || t == nthreads - 1)) h++;
        jobs[t].hi = h;
    }
    jobs[nthreads - 1].hi = n;
    if (nthreads == 1) {
        batch_worker(&jobs[0]);
        return jobs[0].total;
    }
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    uint64_t total = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        total += jobs[t].total;
    }
    return total;
}
