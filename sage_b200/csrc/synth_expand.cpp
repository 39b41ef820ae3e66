// synth_expand.cpp — native helper of sage_b200/synth.py (benchmark / test DATA generation, not part of the search path): enumerates the
// variable-modification forms of a peptide table the way Peptide::apply does (peptide.rs:258-305: the unmodified form, then every combination of
// 1..max_mods modified sites in lexicographic site order), computes each form's monoisotopic mass (f32, sequential: peptide.rs:129-133,
// 361-373), filters by mass and returns the table sorted like reorder_peptides (database.rs:221-258, peptide.rs:34-52: by mass, then sequence
// — the input rows arrive in sequence order — then the modification vectors compared lexicographically). Same order and bits as the numpy implementation in synth.py
// (tests/test_synth.py compares them); ~100x faster, which is what makes the 15 M-peptide table of BASELINE.json's config 3 affordable.
// Build: g++ -O3 -fopenmp -shared -fPIC (sage_b200/build.py: build_synth_library).
#include <parallel/algorithm>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct Expanded {
    std::vector<uint32_t> seq_off;
    std::vector<uint8_t> seq, decoy, missed;
    std::vector<float> mods, mono;
};

// number of combinations of 1..k of n sites (+1 for the unmodified form)
uint64_t forms_of(uint32_t n, uint32_t k) {
    uint64_t total = 1, c = 1;
    for (uint32_t r = 1; r <= k && r <= n; r++) {
        c = c * (n - r + 1) / r;
        total += c;
    }
    return total;
}

}  // namespace

extern "C" void* synth_expand(uint64_t n, uint32_t width, const uint8_t* mat, const int64_t* ln, const float* static_mods, const float* site_mass,
                              const float* base_mono, const uint8_t* decoy, const uint8_t* missed, uint32_t max_mods, float min_mass, float max_mass,
                              uint64_t* n_out, uint64_t* n_res_out) {
    if (max_mods > 4) max_mods = 4;
    // sites of every base peptide
    std::vector<uint64_t> first(n + 1, 0);
    std::vector<uint8_t> nsite(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint32_t c = 0;
        for (int64_t j = 0; j < ln[i]; j++) c += site_mass[(size_t)i * width + j] != 0.0f;
        nsite[i] = (uint8_t)c;
        first[i + 1] = forms_of(c, max_mods);
    }
    for (uint64_t i = 0; i < n; i++) first[i + 1] += first[i];
    const uint64_t nforms = first[n];
    if (nforms >= 0xFFFFFFFFull) return nullptr;
    // form = (base row, up to 4 site columns; 0xFF = unused)
    std::vector<uint32_t> f_base(nforms), f_sites(nforms);
    std::vector<float> f_mono(nforms);
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint8_t cols[64];
        uint32_t ns = 0;
        const size_t row = (size_t)i * width;
        for (int64_t j = 0; j < ln[i] && ns < 64; j++)
            if (site_mass[row + j] != 0.0f) cols[ns++] = (uint8_t)j;
        uint64_t w = first[i];
        auto emit = [&](const uint32_t* idx, uint32_t r) {
            uint32_t packed = 0xFFFFFFFFu;
            for (uint32_t a = 0; a < r; a++) packed = (packed & ~(0xFFu << (8 * a))) | ((uint32_t)cols[idx[a]] << (8 * a));
            // modification_mass: sequential f32 sum over the residues (static mods, variable mods override their sites)
            float ms = 0.0f;
            for (int64_t j = 0; j < ln[i]; j++) {
                float m = static_mods[row + j];
                for (uint32_t a = 0; a < r; a++)
                    if (cols[idx[a]] == j) m = site_mass[row + j];
                ms = ms + m;
            }
            f_base[w] = (uint32_t)i;
            f_sites[w] = packed;
            f_mono[w] = base_mono[i] + ms;
            w++;
        };
        uint32_t idx[4] = {0, 0, 0, 0};
        emit(idx, 0);
        for (uint32_t r = 1; r <= max_mods && r <= ns; r++) {
            for (uint32_t a = 0; a < r; a++) idx[a] = a;
            for (;;) {
                emit(idx, r);
                int a = (int)r - 1;
                while (a >= 0 && idx[a] == ns - r + (uint32_t)a) a--;
                if (a < 0) break;
                idx[a]++;
                for (uint32_t b = (uint32_t)a + 1; b < r; b++) idx[b] = idx[b - 1] + 1;
            }
        }
    }
    // keep [min_mass, max_mass]; order like reorder_peptides: (monoisotopic, sequence == input row, modifications compared lexicographically
    // (peptide.rs:34-52)). Forms of one row with equal mass are positional isomers: at the first column where their site sets differ the form
    // WITHOUT the (positive) modification there is the smaller vector.
    std::vector<uint64_t> keys;
    keys.reserve(nforms);
    for (uint64_t f = 0; f < nforms; f++) {
        const float m = f_mono[f];
        if (m >= min_mass && m <= max_mass) keys.push_back(f);
    }
    auto mod_at = [&](uint32_t f, uint32_t j) -> float {
        const size_t row = (size_t)f_base[f] * width;
        const uint32_t packed = f_sites[f];
        for (uint32_t a = 0; a < 4; a++)
            if (((packed >> (8 * a)) & 0xFFu) == j) return site_mass[row + j];
        return static_mods[row + j];
    };
    __gnu_parallel::sort(keys.begin(), keys.end(), [&](uint64_t a, uint64_t b) {
        const float ma = f_mono[a], mb = f_mono[b];
        if (ma != mb) return ma < mb;
        if (f_base[a] != f_base[b]) return f_base[a] < f_base[b];
        const int64_t L = ln[f_base[a]];
        for (int64_t j = 0; j < L; j++) {
            const float xa = mod_at((uint32_t)a, (uint32_t)j), xb = mod_at((uint32_t)b, (uint32_t)j);
            if (xa != xb) return xa < xb;
        }
        return a < b;
    });
    const uint64_t m_out = keys.size();
    Expanded* E = new Expanded();
    E->seq_off.resize(m_out + 1);
    E->decoy.resize(m_out);
    E->missed.resize(m_out);
    E->mono.resize(m_out);
    uint64_t acc = 0;
    for (uint64_t o = 0; o < m_out; o++) {
        E->seq_off[o] = (uint32_t)acc;
        acc += (uint64_t)ln[f_base[(uint32_t)keys[o]]];
    }
    if (acc > 0xFFFFFFFFull) { delete E; return nullptr; }
    E->seq_off[m_out] = (uint32_t)acc;
    E->seq.resize(acc);
    E->mods.resize(acc);
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < (int64_t)m_out; o++) {
        const uint32_t f = (uint32_t)keys[o], i = f_base[f], packed = f_sites[f];
        const size_t row = (size_t)i * width, dst = E->seq_off[o];
        for (int64_t j = 0; j < ln[i]; j++) {
            E->seq[dst + j] = mat[row + j];
            float m = static_mods[row + j];
            for (uint32_t a = 0; a < 4; a++)
                if (((packed >> (8 * a)) & 0xFFu) == (uint32_t)j) m = site_mass[row + j];
            E->mods[dst + j] = m;
        }
        E->decoy[o] = decoy[i];
        E->missed[o] = missed[i];
        E->mono[o] = f_mono[f];
    }
    *n_out = m_out;
    *n_res_out = acc;
    return E;
}

extern "C" void synth_expand_fetch(void* h, uint32_t* seq_off, uint8_t* seq, float* mods, float* mono, uint8_t* decoy, uint8_t* missed) {
    const Expanded* E = static_cast<const Expanded*>(h);
    memcpy(seq_off, E->seq_off.data(), 4 * E->seq_off.size());
    memcpy(seq, E->seq.data(), E->seq.size());
    memcpy(mods, E->mods.data(), 4 * E->mods.size());
    memcpy(mono, E->mono.data(), 4 * E->mono.size());
    memcpy(decoy, E->decoy.data(), E->decoy.size());
    memcpy(missed, E->missed.data(), E->missed.size());
}

extern "C" void synth_expand_free(void* h) { delete static_cast<Expanded*>(h); }
