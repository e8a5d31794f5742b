/*
 * ms_oracle.c — plain-C restatement of the reference's mutation-injection path.
 *
 * TEST INFRASTRUCTURE — NOT A PRODUCT PATH.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (mutation_simulator_b200) never links or calls it.
 *
 * Parity status: PINNED.  orc_walk/orc_wrap are checked byte-for-byte against
 * golden vectors produced by running the unmodified reference in the build
 * container (tests/golden/, tests/test_oracle_golden.py::test_c_oracle_*).
 * The sampling half uses its own RNG (xoshiro256**); the reference's streams
 * (CPython MT19937 / numpy legacy RandomState) need not be reproduced —
 * fresh-sampling parity is statistical.
 *
 * Every function cites the reference file:line (relative to
 * /root/reference/mutation_simulator/) whose behaviour it restates.
 *
 * Type codes (reference ARGS dict order, rmt.py:443-450 and :91-94):
 *   0 SN, 1 IN, 2 DE, 3 IV, 4 DU, 5 TL, 6 TLI
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { T_SN = 0, T_IN = 1, T_DE = 2, T_IV = 3, T_DU = 4, T_TL = 5, T_TLI = 6 };

typedef struct {
    int64_t key;     /* position at which the walk applies the mutation (dict key) */
    int64_t start;   /* Mutation.start (for TLI: linked TL start) */
    int64_t stop;    /* Mutation.stop  (for TLI: linked TL stop)  */
    int64_t lit_off; /* IN: offset of the insert string in the literal pool */
    int32_t type;
    uint8_t reverse; /* TLI: trans_reverse */
    uint8_t alt;     /* SN: substituted base */
    uint8_t pad[2];
} orc_mut;

typedef struct {
    int64_t start, stop; /* inclusive, 0-based (rmt.py RangeDefinition) */
    int64_t k;           /* int(((stop-start)+1)*sum(rates)), computed by the caller in float64 (mutator.py:225) */
    double cdf[7];       /* cumulative mut_chances (rmt.py:142-151) */
    int32_t minlen[7], maxlen[7];
} orc_range;

/* ---- byte tables: mutator.py:75-77 ------------------------------------- */
static uint8_t NONAMB[256], COMPL[256], TRANS[256];
static int tables_ready = 0;
static void init_tables(void) {
    if (tables_ready) return;
    for (int i = 0; i < 256; i++) NONAMB[i] = COMPL[i] = TRANS[i] = (uint8_t)i;
    const char *a = "KSYMWRBDHV-", *b = "GCCAAACAAAN";
    for (int i = 0; a[i]; i++) NONAMB[(uint8_t)a[i]] = (uint8_t)b[i];
    const char *c = "ACGTUMRWSYKVHDB", *d = "TGCAAKYWSRMBDHV";
    for (int i = 0; c[i]; i++) COMPL[(uint8_t)c[i]] = (uint8_t)d[i];
    const char *e = "AGTC", *f = "GACT";
    for (int i = 0; e[i]; i++) TRANS[(uint8_t)e[i]] = (uint8_t)f[i];
    tables_ready = 1;
}

/* ---- RNG (oracle-private) ---------------------------------------------- */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static void rng_seed(rng_t *r, uint64_t seed) {
    for (int i = 0; i < 4; i++) {
        uint64_t z = (seed += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        r->s[i] = z ^ (z >> 31);
    }
}
static inline double rng_uniform(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
/* uniform integer in [lo, hi], unbiased (rejection), like random.randint */
static int64_t rng_randint(rng_t *r, int64_t lo, int64_t hi) {
    uint64_t n = (uint64_t)(hi - lo) + 1, x, lim = UINT64_MAX - (UINT64_MAX % n + 1) % n;
    do { x = rng_next(r); } while (x > lim);
    return lo + (int64_t)(x % n);
}

/* ---- growable byte sink ------------------------------------------------ */
typedef struct { uint8_t *p; int64_t n, cap; } sink_t;
static int sink_reserve(sink_t *s, int64_t extra) {
    if (s->n + extra <= s->cap) return 0;
    int64_t nc = s->cap ? s->cap : 4096;
    while (nc < s->n + extra) nc *= 2;
    uint8_t *np = (uint8_t *)realloc(s->p, (size_t)nc);
    if (!np) return -1;
    s->p = np; s->cap = nc;
    return 0;
}
static inline void sink_put(sink_t *s, const uint8_t *b, int64_t n) { memcpy(s->p + s->n, b, (size_t)n); s->n += n; }
static void sink_int(sink_t *s, int64_t v) { s->n += sprintf((char *)s->p + s->n, "%lld", (long long)v); }

/* vcf_writer.py:44-52,118-126.  ref/alt are given as up to two pieces each so
 * callers need not concatenate.  Skipped when REF == ALT. */
static int pieces_equal(const uint8_t *a1, int64_t na1, const uint8_t *a2, int64_t na2,
                        const uint8_t *b1, int64_t nb1, const uint8_t *b2, int64_t nb2) {
    if (na1 + na2 != nb1 + nb2) return 0;
    for (int64_t i = 0; i < na1 + na2; i++) {
        uint8_t x = i < na1 ? a1[i] : a2[i - na1], y = i < nb1 ? b1[i] : b2[i - nb1];
        if (x != y) return 0;
    }
    return 1;
}
static int vcf_put(sink_t *v, const char *name, int64_t start, const uint8_t *r1, int64_t nr1,
                   const uint8_t *r2, int64_t nr2, const uint8_t *a1, int64_t na1, const uint8_t *a2,
                   int64_t na2, const char *svtype, int64_t end, int64_t len) {
    if (pieces_equal(r1, nr1, r2, nr2, a1, na1, a2, na2)) return 0;
    int64_t nl = (int64_t)strlen(name);
    if (sink_reserve(v, nl + nr1 + nr2 + na1 + na2 + 128)) return -1;
    sink_put(v, (const uint8_t *)name, nl);
    v->p[v->n++] = '\t'; sink_int(v, start);
    sink_put(v, (const uint8_t *)"\t.\t", 3);
    sink_put(v, r1, nr1); sink_put(v, r2, nr2);
    v->p[v->n++] = '\t';
    sink_put(v, a1, na1); sink_put(v, a2, na2);
    sink_put(v, (const uint8_t *)"\t.\t.\t", 5);
    if (!svtype) v->p[v->n++] = '.';
    else {
        sink_put(v, (const uint8_t *)"SVTYPE=", 7); sink_put(v, (const uint8_t *)svtype, (int64_t)strlen(svtype));
        sink_put(v, (const uint8_t *)";END=", 5); sink_int(v, end);
        sink_put(v, (const uint8_t *)";SVLEN=", 7); sink_int(v, len);
    }
    sink_put(v, (const uint8_t *)"\tGT\t1\n", 6);
    return 0;
}

static void conv_copy(uint8_t *dst, const uint8_t *src, int64_t n) { for (int64_t i = 0; i < n; i++) dst[i] = NONAMB[src[i]]; }
static void revcomp_conv(uint8_t *dst, const uint8_t *src, int64_t n) { /* [::-1].translate(complement) of conv(src) */
    for (int64_t i = 0; i < n; i++) dst[i] = COMPL[NONAMB[src[n - 1 - i]]];
}

/*
 * mutator.py:318-426 __mutate_sequence.  seq is the upper-cased contig; muts are
 * sorted by key.  Mutations whose key lies inside an earlier DE/TL/IV/DU span
 * are skipped, as the reference's `pos` jump does (:376,386,398).
 * body/vcf are malloc'd sinks owned by the caller (free with orc_free).
 */
int orc_walk(const uint8_t *seq, int64_t L, const char *name, const orc_mut *muts, int64_t n,
             const uint8_t *lit, uint8_t **body_out, int64_t *body_len, uint8_t **vcf_out, int64_t *vcf_len) {
    init_tables();
    sink_t b = {0, 0, 0}, v = {0, 0, 0};
    if (sink_reserve(&b, L + 64) || sink_reserve(&v, 64)) return -1;
    uint8_t *tmp = NULL; int64_t tmpcap = 0;
    int64_t pos = 0;
    for (int64_t i = 0; i < n; i++) {
        const orc_mut *m = &muts[i];
        int64_t p = m->key;
        if (p < pos || p >= L) continue;
        if (sink_reserve(&b, p - pos)) return -1;
        sink_put(&b, seq + pos, p - pos);
        int64_t len = m->stop - m->start + 1;
        int64_t need = 2 * (len > 0 ? len : 0) + 16;
        if (need > tmpcap) { tmp = (uint8_t *)realloc(tmp, (size_t)need); tmpcap = need; if (!tmp) return -1; }
        switch (m->type) {
        case T_SN: { /* :334-341 */
            uint8_t ref = NONAMB[seq[p]], alt = m->alt;
            if (sink_reserve(&b, 1)) return -1;
            b.p[b.n++] = alt;
            if (vcf_put(&v, name, p + 1, &ref, 1, NULL, 0, &alt, 1, NULL, 0, NULL, 0, 0)) return -1;
            pos = p + 1;
        } break;
        case T_IN: { /* :343-358 */
            const uint8_t *ins = lit + m->lit_off;
            uint8_t ref = NONAMB[seq[p > 0 ? p - 1 : 0]];
            if (p > 0) { if (vcf_put(&v, name, p, &ref, 1, NULL, 0, &ref, 1, ins, len, "INS", p, len)) return -1; }
            else       { if (vcf_put(&v, name, 1, &ref, 1, NULL, 0, ins, len, &ref, 1, "INS", 1, len)) return -1; }
            if (sink_reserve(&b, len + 1)) return -1;
            sink_put(&b, ins, len); b.p[b.n++] = seq[p];
            pos = p + 1;
        } break;
        case T_DE: case T_TL: { /* :360-377 */
            const char *sv = m->type == T_DE ? "DEL" : "DEL:ME";
            int64_t end = m->stop + 1;
            if (p > 0) {
                int64_t rl = end - (p - 1);
                if (rl + 1 > tmpcap) { tmp = (uint8_t *)realloc(tmp, (size_t)rl + 16); tmpcap = rl + 16; }
                conv_copy(tmp, seq + p - 1, rl);
                if (vcf_put(&v, name, p, tmp, rl, NULL, 0, tmp, 1, NULL, 0, sv, end, m->stop - p + 1)) return -1;
            } else {
                end += 1;
                int64_t rl = end > L ? L : end;
                if (rl + 1 > tmpcap) { tmp = (uint8_t *)realloc(tmp, (size_t)rl + 16); tmpcap = rl + 16; }
                conv_copy(tmp, seq, rl);
                if (vcf_put(&v, name, 1, tmp, rl, NULL, 0, tmp + rl - 1, 1, NULL, 0, sv, end, m->stop - p + 1)) return -1;
            }
            pos = m->stop + 1;
        } break;
        case T_IV: { /* :379-387 */
            conv_copy(tmp, seq + p, len);
            revcomp_conv(tmp + len, seq + p, len);
            if (sink_reserve(&b, len)) return -1;
            sink_put(&b, tmp + len, len);
            if (vcf_put(&v, name, p + 1, tmp, len, NULL, 0, tmp + len, len, NULL, 0, "INV", m->stop + 1, 0)) return -1;
            pos = m->stop + 1;
        } break;
        case T_DU: { /* :389-399 (raw bases, not converted) */
            if (sink_reserve(&b, 2 * len)) return -1;
            sink_put(&b, seq + p, len); sink_put(&b, seq + p, len);
            if (vcf_put(&v, name, p + 1, seq + p, len, NULL, 0, seq + p, len, seq + p, len, "DUP", p + len, len)) return -1;
            pos = m->stop + 1;
        } break;
        case T_TLI: { /* :401-421 */
            if (m->reverse) revcomp_conv(tmp, seq + m->start, len); else conv_copy(tmp, seq + m->start, len);
            uint8_t ref = NONAMB[seq[p > 0 ? p - 1 : p]];
            if (p > 0) { if (vcf_put(&v, name, p, &ref, 1, NULL, 0, &ref, 1, tmp, len, "INS:ME", p, len)) return -1; }
            else       { if (vcf_put(&v, name, 1, &ref, 1, NULL, 0, tmp, len, &ref, 1, "INS:ME", 1, len)) return -1; }
            if (sink_reserve(&b, len + 1)) return -1;
            sink_put(&b, tmp, len); b.p[b.n++] = seq[p];
            pos = p + 1;
        } break;
        default: free(tmp); return -2;
        }
    }
    if (sink_reserve(&b, L - pos)) return -1;
    sink_put(&b, seq + pos, L - pos);
    free(tmp);
    *body_out = b.p; *body_len = b.n; *vcf_out = v.p; *vcf_len = v.n;
    return 0;
}

void orc_free(void *p) { free(p); }

/*
 * fasta_writer.py:40-65 for one contig: ">header\n" then the body with a line
 * break after every bpl bases.  *written carries FastaWriter.__written between
 * contigs (a header is preceded by "\n" only if the previous line is partial).
 * Returns bytes written to dst (dst must hold 2+hlen+1+n+n/bpl+1).
 */
int64_t orc_wrap(const uint8_t *header, int64_t hlen, const uint8_t *body, int64_t n, int64_t bpl,
                 int64_t *written, uint8_t *dst) {
    int64_t o = 0;
    if (*written != 0) dst[o++] = '\n';
    dst[o++] = '>'; memcpy(dst + o, header, (size_t)hlen); o += hlen; dst[o++] = '\n';
    int64_t i = 0;
    while (i + bpl <= n) { memcpy(dst + o, body + i, (size_t)bpl); o += bpl; dst[o++] = '\n'; i += bpl; }
    memcpy(dst + o, body + i, (size_t)(n - i)); o += n - i;
    *written = n - i;
    return o;
}

/* ---- sampling half ------------------------------------------------------ */
static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y); }

/* util.py:94-109 sample_with_minimum_distance: uniform k-subset of
 * range(start, stop-(k-1)d), sorted, + d*rank.  Returns -1 like random.sample's
 * ValueError when k > population or k < 0. */
static int64_t sample_min_dist(rng_t *r, int64_t start, int64_t stop, int64_t k, int64_t d, int64_t *out) {
    int64_t n = (stop - (k - 1) * d) - start;
    if (k < 0 || k > n) return -1;
    if (k == 0) return 0;
    if (k * 3 > n) { /* dense: partial Fisher-Yates over the population */
        int64_t *pool = (int64_t *)malloc((size_t)n * sizeof(int64_t));
        if (!pool) return -1;
        for (int64_t i = 0; i < n; i++) pool[i] = i;
        for (int64_t i = 0; i < k; i++) { int64_t j = rng_randint(r, i, n - 1), t = pool[i]; pool[i] = pool[j]; pool[j] = t; out[i] = pool[i]; }
        free(pool);
    } else { /* sparse: draw until distinct (what CPython's sample does with a set) */
        uint64_t cap = 16; while (cap < (uint64_t)k * 3) cap <<= 1;
        int64_t *tab = (int64_t *)malloc(cap * sizeof(int64_t));
        if (!tab) return -1;
        memset(tab, 0xff, cap * sizeof(int64_t));
        for (int64_t i = 0; i < k;) {
            int64_t v = rng_randint(r, 0, n - 1);
            uint64_t h = ((uint64_t)v * 0x9e3779b97f4a7c15ULL) & (cap - 1);
            while (tab[h] != -1 && tab[h] != v) h = (h + 1) & (cap - 1);
            if (tab[h] == v) continue;
            tab[h] = v; out[i++] = v;
        }
        free(tab);
    }
    qsort(out, (size_t)k, sizeof(int64_t), cmp_i64);
    for (int64_t i = 0; i < k; i++) out[i] += start + d * i;
    return k;
}

/*
 * mutator.py:144-316: per range sample positions (:217-226), assign types
 * (:166-182), greedy first-come rejection with stop assignment (:184-213,
 * :229-265), then link TL<->TLI across the contig (:268-316).  SN alt bases
 * (:429-463) and insert strings (:466-471) are drawn here as well, so that
 * orc_walk is deterministic.  Reference semantics (blocking state resets per
 * range, no clipping at blocked ranges) are kept — this is the restatement.
 * Returns number of mutations (sorted by key) or <0 on error.
 */
int64_t orc_sample_contig(const uint8_t *seq, int64_t L, const orc_range *ranges, int32_t n_ranges,
                          const int32_t block[7], int32_t min_dist, double titv, uint64_t seed,
                          orc_mut **muts_out, uint8_t **lit_out, int64_t *lit_len) {
    init_tables();
    rng_t r; rng_seed(&r, seed);
    int64_t ktot = 0;
    for (int i = 0; i < n_ranges; i++) ktot += ranges[i].k > 0 ? ranges[i].k : 0;
    orc_mut *m = (orc_mut *)malloc(((size_t)ktot + 1) * sizeof(orc_mut));
    int64_t *pos = (int64_t *)malloc(((size_t)ktot + 1) * sizeof(int64_t));
    if (!m || !pos) return -1;
    int64_t nm = 0;
    for (int ri = 0; ri < n_ranges; ri++) {
        const orc_range *rg = &ranges[ri];
        if (rg->k == 0) continue;
        int64_t k = sample_min_dist(&r, rg->start, rg->stop, rg->k, min_dist, pos);
        if (k < 0) { free(m); free(pos); return -3; }
        int64_t last_hi = -1;
        for (int64_t i = 0; i < k; i++) {
            double u = rng_uniform(&r);
            int t = 0; while (t < 6 && u >= rg->cdf[t]) t++;
            int64_t p = pos[i];
            if (p < last_hi) continue;               /* pos in last_mut_range (:190) */
            int64_t stop;
            switch (t) {                             /* :229-265 */
            case T_SN: stop = p; break;
            case T_IV:
                if (p + rg->maxlen[T_IV] >= L - 1) continue;
                stop = rng_randint(&r, p + rg->minlen[T_IV] - 1, p + rg->maxlen[T_IV] - 1); break;
            case T_IN: stop = rng_randint(&r, p + rg->minlen[T_IN] - 1, p + rg->maxlen[T_IN] - 1); break;
            case T_DU: case T_DE: case T_TL:
                stop = rng_randint(&r, p + rg->minlen[t] - 1, p + rg->maxlen[t] - 1);
                if (stop > L - 1) stop = L - 1; break;
            default: stop = 0; break;                /* TLI placeholder */
            }
            orc_mut *q = &m[nm++];
            memset(q, 0, sizeof(*q));
            q->key = p; q->start = p; q->stop = stop; q->type = t;
            last_hi = ((t == T_SN || t == T_IN) ? p : stop) + 1 + block[t];   /* :204-209 */
        }
    }
    /* __link_tls / __fix_tl_amount (:268-304): random surplus removal, shuffle TLs, zip with TLIs in order */
    int64_t ntl = 0, ntli = 0;
    for (int64_t i = 0; i < nm; i++) { ntl += m[i].type == T_TL; ntli += m[i].type == T_TLI; }
    if (ntl > 0) {
        int64_t *tl = (int64_t *)malloc((size_t)(ntl + 1) * sizeof(int64_t)), *tli = (int64_t *)malloc((size_t)(ntli + 1) * sizeof(int64_t));
        int64_t a = 0, b = 0;
        for (int64_t i = 0; i < nm; i++) { if (m[i].type == T_TL) tl[a++] = i; else if (m[i].type == T_TLI) tli[b++] = i; }
        while (a < b) { int64_t j = rng_randint(&r, 0, b - 1); m[tli[j]].type = -1; memmove(tli + j, tli + j + 1, (size_t)(b - j - 1) * sizeof(int64_t)); b--; }
        while (a > b) { int64_t j = rng_randint(&r, 0, a - 1); m[tl[j]].type = -1; memmove(tl + j, tl + j + 1, (size_t)(a - j - 1) * sizeof(int64_t)); a--; }
        for (int64_t i = a - 1; i > 0; i--) { int64_t j = rng_randint(&r, 0, i), t = tl[i]; tl[i] = tl[j]; tl[j] = t; }
        for (int64_t i = 0; i < a; i++) {
            orc_mut *s = &m[tl[i]], *d = &m[tli[i]];
            d->start = s->start; d->stop = s->stop;
            int64_t len = s->stop + 1 - s->start;
            d->reverse = !(rng_randint(&r, 0, 1) == 0 || len < 2);   /* :307-316 */
        }
        free(tl); free(tli);
    } else {
        /* no TL at all: unlinked TLI placeholders stay in the dict but emit nothing for key>0; drop them */
        for (int64_t i = 0; i < nm; i++) if (m[i].type == T_TLI) m[i].type = -1;
    }
    int64_t w = 0, litn = 0;
    for (int64_t i = 0; i < nm; i++) if (m[i].type >= 0) { m[w++] = m[i]; if (m[w - 1].type == T_IN) litn += m[w - 1].stop - m[w - 1].start + 1; }
    nm = w;
    uint8_t *lit = (uint8_t *)malloc((size_t)litn + 1);
    int64_t lo = 0;
    double p_ti = titv * (1 / (titv + 1));
    static const char ATGC[4] = {'A', 'T', 'G', 'C'};
    for (int64_t i = 0; i < nm; i++) {
        if (m[i].type == T_SN) {
            uint8_t ref = NONAMB[seq[m[i].key]];
            if (rng_uniform(&r) <= p_ti) m[i].alt = TRANS[ref];
            else {
                int c = (int)rng_randint(&r, 0, 1);
                switch (ref) {
                case 'A': m[i].alt = "TC"[c]; break; case 'G': m[i].alt = "CT"[c]; break;
                case 'T': m[i].alt = "GA"[c]; break; case 'C': m[i].alt = "AG"[c]; break;
                default: m[i].alt = ref;
                }
            }
        } else if (m[i].type == T_IN) {
            int64_t len = m[i].stop - m[i].start + 1;
            m[i].lit_off = lo;
            for (int64_t j = 0; j < len; j++) lit[lo++] = (uint8_t)ATGC[rng_next(&r) >> 62];
        }
    }
    free(pos);
    *muts_out = m; *lit_out = lit; *lit_len = litn;
    return nm;
}

/*
 * One contig end to end, as Mutator.mutate does per chromosome (mutator.py:111-141):
 * sample -> link -> walk -> wrapped FASTA + VCF.  Used for the CPU baseline timing
 * and for statistics.  Outputs are malloc'd (free with orc_free).
 */
int64_t orc_mutate_contig(const uint8_t *seq, int64_t L, const char *name, const uint8_t *header, int64_t hlen,
                          int64_t bpl, const orc_range *ranges, int32_t n_ranges, const int32_t block[7],
                          int32_t min_dist, double titv, uint64_t seed, int64_t *written,
                          uint8_t **fasta_out, int64_t *fasta_len, uint8_t **vcf_out, int64_t *vcf_len,
                          int64_t counts[7]) {
    orc_mut *m = NULL; uint8_t *lit = NULL; int64_t litn = 0;
    int64_t nm = orc_sample_contig(seq, L, ranges, n_ranges, block, min_dist, titv, seed, &m, &lit, &litn);
    if (nm < 0) return nm;
    if (counts) { memset(counts, 0, 7 * sizeof(int64_t)); for (int64_t i = 0; i < nm; i++) counts[m[i].type]++; }
    uint8_t *body = NULL; int64_t bl = 0;
    int rc = orc_walk(seq, L, name, m, nm, lit, &body, &bl, vcf_out, vcf_len);
    free(m); free(lit);
    if (rc) return rc;
    uint8_t *fa = (uint8_t *)malloc((size_t)(bl + bl / (bpl > 0 ? bpl : 1) + hlen + 8));
    if (!fa) return -1;
    *fasta_len = orc_wrap(header, hlen, body, bl, bpl, written, fa);
    *fasta_out = fa;
    free(body);
    return nm;
}
