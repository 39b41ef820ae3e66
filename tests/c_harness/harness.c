/* harness.c — drives libsage_b200.so through include/sage_b200.h from plain C, the way the Rust shim of INTEGRATION.md does:
 * array-of-structs inputs (one `peptide` / `spectrum` record each, as sage-core holds Vec<Peptide> / Vec<ProcessedSpectrum>) are flattened
 * into the SoA views of the C ABI in malloc'd (pageable) memory, ONE scorer handle is kept for the life of the device database and reused for
 * every batch, errors are fetched with sage_b200_last_error, and annotate_matches' Fragments are re-attached per PSM.
 *
 *   harness <in.bin> <out.bin>          (tests/test_c_harness.py writes in.bin, reads out.bin and compares with the oracle)
 *
 * in.bin  : u64 n_pep | per peptide {u32 len, u8 decoy, u8 missed, f32 mono, f32 nterm, u8 seq[len], f32 mods[len]}
 *           | scorer params (sage_b200_scorer_params, raw) | u64 n_batches | per batch: u64 n_spec | per spectrum {u32 n_peaks, f32 prec_mz,
 *           u8 charge, u8 level, f32 iso_lo, f32 iso_hi, f32 tic, f32 rt, f32 ims, f32 mass[n_peaks], f32 intensity[n_peaks]}
 * out.bin : per batch: i32 rc | (rc == 0: u64 n_spec, u32 counts[n_spec], sage_b200_feature feats[n_spec * report_psms], u64 n_frag,
 *           sage_b200_fragment frags[n_frag]) | (rc != 0: u32 len, char msg[len])
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sage_b200.h"

typedef struct { uint32_t len; uint8_t decoy, missed; float mono, nterm; uint8_t* seq; float* mods; } peptide;      /* ~ peptide.rs:13-31 */
typedef struct { uint32_t n_peaks; float prec_mz; uint8_t charge, level; float iso_lo, iso_hi, tic, rt, ims; float *mass, *intensity; } spectrum; /* ~ spectrum.rs:58-79 */

static void die(const char* what) {
    char buf[1024];
    sage_b200_last_error(buf, sizeof buf);
    fprintf(stderr, "harness: %s: %s\n", what, buf);
    exit(2);
}
static void rd(void* p, size_t n, FILE* f) { if (n && fread(p, 1, n, f) != n) { fprintf(stderr, "harness: short read\n"); exit(3); } }

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    FILE* in = fopen(argv[1], "rb");
    FILE* out = fopen(argv[2], "wb");
    if (!in || !out) return 1;
    /* ---- Vec<Peptide> */
    uint64_t n_pep;
    rd(&n_pep, 8, in);
    peptide* peps = calloc(n_pep ? n_pep : 1, sizeof *peps);
    uint64_t n_res = 0;
    for (uint64_t i = 0; i < n_pep; i++) {
        peptide* p = &peps[i];
        rd(&p->len, 4, in); rd(&p->decoy, 1, in); rd(&p->missed, 1, in); rd(&p->mono, 4, in); rd(&p->nterm, 4, in);
        p->seq = malloc(p->len); p->mods = malloc(4 * (size_t)p->len);
        rd(p->seq, p->len, in); rd(p->mods, 4 * (size_t)p->len, in);
        n_res += p->len;
    }
    /* flatten (DeviceDatabase::upload of INTEGRATION.md) */
    uint32_t* off = malloc(4 * (n_pep + 1));
    uint8_t *seq = malloc(n_res + 1), *decoy = malloc(n_pep + 1), *missed = malloc(n_pep + 1);
    float *mods = malloc(4 * (n_res + 1)), *mono = malloc(4 * (n_pep + 1)), *nterm = malloc(4 * (n_pep + 1));
    uint64_t w = 0;
    for (uint64_t i = 0; i < n_pep; i++) {
        off[i] = (uint32_t)w;
        memcpy(seq + w, peps[i].seq, peps[i].len); memcpy(mods + w, peps[i].mods, 4 * (size_t)peps[i].len);
        w += peps[i].len;
        mono[i] = peps[i].mono; nterm[i] = peps[i].nterm; decoy[i] = peps[i].decoy; missed[i] = peps[i].missed;
    }
    off[n_pep] = (uint32_t)w;
    sage_b200_peptides P = {n_pep, off, seq, mods, nterm, mono, decoy, missed};
    const uint8_t kinds[2] = {SAGE_B200_KIND_B, SAGE_B200_KIND_Y};
    sage_b200_db* db = NULL;
    if (sage_b200_db_build(&P, 8192, kinds, 2, 2, 0, &db)) die("db_build");
    sage_b200_scorer_params sp;
    rd(&sp, sizeof sp, in);
    sage_b200_scorer* scorer = NULL;                       /* ONE handle for every batch: it keeps its learned work-list sizes */
    if (sage_b200_scorer_create(db, &sp, &scorer)) die("scorer_create");
    uint64_t n_batches;
    rd(&n_batches, 8, in);
    for (uint64_t b = 0; b < n_batches; b++) {
        uint64_t n;
        rd(&n, 8, in);
        spectrum* S = calloc(n ? n : 1, sizeof *S);
        uint64_t n_peaks = 0;
        for (uint64_t i = 0; i < n; i++) {
            spectrum* s = &S[i];
            rd(&s->n_peaks, 4, in); rd(&s->prec_mz, 4, in); rd(&s->charge, 1, in); rd(&s->level, 1, in); rd(&s->iso_lo, 4, in); rd(&s->iso_hi, 4, in);
            rd(&s->tic, 4, in); rd(&s->rt, 4, in); rd(&s->ims, 4, in);
            s->mass = malloc(4 * (size_t)s->n_peaks + 4); s->intensity = malloc(4 * (size_t)s->n_peaks + 4);
            rd(s->mass, 4 * (size_t)s->n_peaks, in); rd(s->intensity, 4 * (size_t)s->n_peaks, in);
            n_peaks += s->n_peaks;
        }
        /* SoA flattening of &[ProcessedSpectrum] into pageable memory (Scorer::score_batch of INTEGRATION.md) */
        uint64_t* poff = malloc(8 * (n + 1));
        float *m = malloc(4 * (n_peaks + 1)), *it = malloc(4 * (n_peaks + 1)), *pmz = malloc(4 * (n + 1)), *ilo = malloc(4 * (n + 1)), *ihi = malloc(4 * (n + 1)),
              *tic = malloc(4 * (n + 1)), *rt = malloc(4 * (n + 1)), *ims = malloc(4 * (n + 1));
        uint8_t *chg = malloc(n + 1), *lvl = malloc(n + 1);
        uint64_t pw = 0;
        for (uint64_t i = 0; i < n; i++) {
            poff[i] = pw;
            memcpy(m + pw, S[i].mass, 4 * (size_t)S[i].n_peaks); memcpy(it + pw, S[i].intensity, 4 * (size_t)S[i].n_peaks);
            pw += S[i].n_peaks;
            pmz[i] = S[i].prec_mz; chg[i] = S[i].charge; lvl[i] = S[i].level; ilo[i] = S[i].iso_lo; ihi[i] = S[i].iso_hi; tic[i] = S[i].tic; rt[i] = S[i].rt; ims[i] = S[i].ims;
        }
        poff[n] = pw;
        sage_b200_spectra V = {n, poff, m, it, pmz, chg, ilo, ihi, tic, lvl, rt, ims};
        const uint64_t nf = n * sp.report_psms;
        sage_b200_feature* feats = calloc(nf ? nf : 1, sizeof *feats);
        uint32_t* counts = calloc(n ? n : 1, 4);
        const uint64_t fcap = sp.annotate_matches ? nf * 128 + 1024 : 0;
        sage_b200_fragment* frags = sp.annotate_matches ? calloc(fcap, sizeof *frags) : NULL;
        uint64_t fused = 0;
        const int32_t rc = sage_b200_score_batch(scorer, &V, feats, counts, frags, fcap, &fused);
        fwrite(&rc, 4, 1, out);
        if (rc == 0) {
            fwrite(&n, 8, 1, out); fwrite(counts, 4, n, out); fwrite(feats, sizeof *feats, nf, out);
            fwrite(&fused, 8, 1, out);
            if (fused) fwrite(frags, sizeof *frags, fused, out);
        } else {   /* the shim turns these into the reference's panics (scoring.rs:301-304, 466-468) */
            char buf[1024];
            const uint32_t len = (uint32_t)sage_b200_last_error(buf, sizeof buf);
            const uint32_t wl = len < sizeof buf ? len : (uint32_t)sizeof buf - 1;
            fwrite(&wl, 4, 1, out); fwrite(buf, 1, wl, out);
        }
        for (uint64_t i = 0; i < n; i++) { free(S[i].mass); free(S[i].intensity); }
        free(S); free(poff); free(m); free(it); free(pmz); free(ilo); free(ihi); free(tic); free(rt); free(ims); free(chg); free(lvl); free(feats); free(counts); free(frags);
    }
    sage_b200_scorer_destroy(scorer);
    sage_b200_db_destroy(db);
    fclose(in); fclose(out);
    return 0;
}
