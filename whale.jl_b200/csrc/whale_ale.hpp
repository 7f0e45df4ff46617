// Native `.ale` ingest: ALEobserve text files -> the reference's CCD layout (src/ccd.jl:126-248), parsed on all
// host threads straight into the flattened whale_ccd_desc arrays that whale_data_create packs for the device.
// The reference does this per file under `tmap` (src/ccd.jl:134); semantics follow it section by section:
//   parse      `#`-separated sections constructor_string, observations, Bip_counts, Bip_bls, Dip_counts,
//              last_leafset_id, leaf-id, set-id, END                                        (:147-196)
//   addleafclades!  leaf clades get count = observations, leaf ids are remapped to set ids   (:199-216)
//   addubiquitous!  clade Γ = all complementary pairs (i < j), counts summed                  (:219-248)
//   CCD(...)   new ids by (size, old id), p = count / count(parent), triples in file order,
//              compat[e] = ascending clades whose species ⊆ clade(e)                         (:102-121)
// Host code only (no device work here).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace whale_ale {

struct Family {  // one CCD in the reference layout (family-local ids)
    std::vector<int32_t> nleaf;
    std::vector<int64_t> split_off;  // [Γ+1]
    std::vector<int32_t> g1, g2;
    std::vector<double> p;
    std::vector<int64_t> compat_off;  // [nn+1]
    std::vector<int32_t> compat;
    std::string error;
};

struct Species {  // what the species tree contributes: gene-name prefix -> species id, species sets per node
    std::unordered_map<std::string, int> id_of;
    int n_ids = 0;                           // ids are 1..n_ids
    std::vector<std::vector<uint64_t>> node_mask;  // [nn] bitset over species ids
    int words = 1;
};

// whitespace-separated tokens of the line [b, e) as (pointer, length) pairs into the section's own text: no copies.
// A token is followed by whitespace or by the NUL that ends the section's std::string, so strtol / strtod may be called on
// its pointer directly (they stop there, like atoi / atof on a copy of the token).
struct Tok { const char* p; size_t n; };
static inline void split_ws(const char* b, const char* e, std::vector<Tok>& out) {
    out.clear();
    const char* i = b;
    while (i < e) {
        while (i < e && (*i == ' ' || *i == '\t' || *i == '\r')) i++;
        const char* j = i;
        while (j < e && *j != ' ' && *j != '\t' && *j != '\r') j++;
        if (j > i) out.push_back(Tok{i, (size_t)(j - i)});
        i = j;
    }
}
// plain decimal integers (what ALEobserve writes for ids and counts) are converted in place; anything else goes to
// strtol / strtod, so the result is what atoi / atof give on the token
static inline int tok_int(const Tok& t) {
    if (t.n >= 1 && t.n <= 9) {
        int v = 0;
        size_t i = 0;
        for (; i < t.n && t.p[i] >= '0' && t.p[i] <= '9'; i++) v = v * 10 + (t.p[i] - '0');
        if (i == t.n) return v;
    }
    return (int)strtol(t.p, nullptr, 10);
}
static inline double tok_dbl(const Tok& t) {
    if (t.n >= 1 && t.n <= 15) {  // up to 15 digits: exactly representable, the conversion is exact like strtod's
        long long v = 0;
        size_t i = 0;
        for (; i < t.n && t.p[i] >= '0' && t.p[i] <= '9'; i++) v = v * 10 + (t.p[i] - '0');
        if (i == t.n) return (double)v;
    }
    return strtod(t.p, nullptr);
}

static inline void parse_family(const std::string& path, const Species& sp, int nn, Family& F) {
    FILE* fh = fopen(path.c_str(), "rb");
    if (!fh) throw std::runtime_error("cannot open " + path);
    std::string text;
    {
        long sz = -1;
        if (fseek(fh, 0, SEEK_END) == 0) { sz = ftell(fh); rewind(fh); }
        if (sz > 0) {
            text.resize((size_t)sz);
            const size_t got = fread(&text[0], 1, (size_t)sz, fh);
            text.resize(got);
        }
        char buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), fh)) > 0) text.append(buf, n);  // (not seekable, or grown meanwhile)
    }
    fclose(fh);
    // `#`-separated parts; inside a part every ":\t" is dropped (repeatedly, like find/erase from the start would)
    std::vector<std::string> parts;
    {
        size_t a = 0;
        for (;;) {
            size_t b = text.find('#', a);
            const size_t e = b == std::string::npos ? text.size() : b;
            parts.emplace_back();
            std::string& out = parts.back();
            out.reserve(e - a);
            for (size_t k = a; k < e;) {  // copy runs up to the next tab; a tab right behind a ':' takes the ':' with it
                const char* tab = static_cast<const char*>(memchr(text.data() + k, '\t', e - k));
                const size_t stop = tab ? (size_t)(tab - text.data()) : e;
                out.append(text, k, stop - k);
                if (stop == e) break;
                if (!out.empty() && out.back() == ':') out.pop_back();
                else out.push_back('\t');
                k = stop + 1;
            }
            if (b == std::string::npos) break;
            a = b + 1;
        }
    }
    if (parts.size() != 10) throw std::runtime_error("Not a valid .ale file " + path);
    // sections: name (first non-empty line, trailing blanks dropped, '-' -> '_') -> its other non-empty lines as ranges
    struct Sec { const std::string* part; std::vector<std::pair<size_t, size_t>> lines; };
    std::map<std::string, Sec> sec;
    for (size_t pi = 1; pi + 1 < parts.size(); pi++) {
        const std::string& part = parts[pi];
        std::string name;
        bool have_name = false;
        Sec cur{&part, {}};
        size_t a = 0;
        while (a <= part.size()) {
            size_t b = part.find('\n', a);
            if (b == std::string::npos) b = part.size();
            size_t e = b;
            if (e > a && part[e - 1] == '\r') e--;
            if (e > a) {
                if (!have_name) {
                    name = part.substr(a, e - a);
                    while (!name.empty() && (name.back() == ' ' || name.back() == '\t')) name.pop_back();
                    std::replace(name.begin(), name.end(), '-', '_');
                    have_name = true;
                } else {
                    cur.lines.emplace_back(a, e);
                }
            }
            a = b + 1;
        }
        if (!have_name) continue;
        sec[name] = std::move(cur);
    }
    auto need = [&](const char* k) -> const Sec& {
        auto it = sec.find(k);
        if (it == sec.end()) throw std::runtime_error(std::string("section ") + k + " missing in " + path);
        return it->second;
    };
    std::vector<Tok> t;
    auto toks = [&](const Sec& S, size_t li) { split_ws(S.part->data() + S.lines[li].first, S.part->data() + S.lines[li].second, t); };
    double obs;
    {
        const Sec& S = need("observations");
        if (S.lines.empty()) throw std::out_of_range("observations section is empty in " + path);
        obs = strtod(S.part->data() + S.lines[0].first, nullptr);
    }
    std::map<int, double> bip;
    {
        const Sec& S = need("Bip_counts");
        for (size_t li = 0; li < S.lines.size(); li++) { toks(S, li); if (t.size() >= 2) bip[tok_int(t[0])] = tok_dbl(t[1]); }
    }
    struct Trip { int a, b; double c; };
    std::map<int, std::vector<Trip>> dip;
    {
        const Sec& S = need("Dip_counts");
        for (size_t li = 0; li < S.lines.size(); li++) {
            toks(S, li);
            if (t.size() < 4) continue;
            dip[tok_int(t[0])].push_back(Trip{tok_int(t[1]), tok_int(t[2]), tok_dbl(t[3])});
        }
    }
    std::map<int, std::string> leaf_name;  // leaf id -> gene name
    {
        const Sec& S = need("leaf_id");
        for (size_t li = 0; li < S.lines.size(); li++) { toks(S, li); if (t.size() >= 2) leaf_name[tok_int(t[1])] = std::string(t[0].p, t[0].n); }
    }
    std::map<int, std::vector<int>> sets;
    {
        const Sec& S = need("set_id");
        for (size_t li = 0; li < S.lines.size(); li++) {
            toks(S, li);
            if (t.empty()) continue;
            std::vector<int> v;
            v.reserve(t.size() - 1);
            for (size_t i = 1; i < t.size(); i++) v.push_back(tok_int(t[i]));
            sets[tok_int(t[0])] = std::move(v);
        }
    }
    // addleafclades!
    int max_leaf = 0;
    for (auto& kv : leaf_name) max_leaf = std::max(max_leaf, kv.first);
    std::vector<int> leaf2set((size_t)max_leaf + 1, 0);  // leaf id -> set id of its leaf clade (0: none; set ids start at 1)
    std::map<int, std::string> leafname;  // set id of a leaf clade -> gene name
    for (auto& kv : sets) {
        const int k = kv.first;
        std::vector<int>& v = kv.second;
        if (v.size() == 1) {
            bip[k] = obs;
            dip[k].clear();
            auto it = leaf_name.find(v[0]);
            if (it == leaf_name.end()) throw std::runtime_error("leaf id without a name in " + path);
            leafname[k] = it->second;
            leaf2set[v[0]] = k;
        } else {
            for (int& i : v) {
                if (i < 0 || i > max_leaf || leaf2set[i] == 0) throw std::runtime_error("set refers to an unknown leaf in " + path);
                i = leaf2set[i];
            }
        }
    }
    for (int l = 0; l <= max_leaf; l++) if (leaf2set[l] != 0) sets[leaf2set[l]] = std::vector<int>{leaf2set[l]};
    // addubiquitous!
    const int nleaves = (int)leafname.size();
    const int ns = (int)sets.size();
    const int G = ns + 1;
    if (G > 65535) throw std::runtime_error("more than 65535 clades (UInt16 ids) in " + path);
    const int W = (ns + 64) / 64;
    std::vector<uint64_t> fs((size_t)(ns + 1) * W, 0);  // clade k as a bitset over set ids: fs[k*W ..]
    std::vector<int> fsz(ns + 1, 0);
    std::vector<double> bipv(G + 1, 0.0);               // counts by id (what the map would default-construct: 0)
    for (auto& kv : bip) if (kv.first >= 0 && kv.first <= G) bipv[kv.first] = kv.second;
    {
        int k = 0;
        for (auto& kv : sets) {  // ascending keys: must be exactly 1..ns
            if (kv.first != ++k) throw std::runtime_error("set ids are not 1..n in " + path);
            uint64_t* row = fs.data() + (size_t)k * W;
            for (int g : kv.second) {
                if (g < 0 || g > ns) throw std::runtime_error("set ids are not 1..n in " + path);
                row[g >> 6] |= 1ull << (g & 63);
            }
            int c = 0;
            for (int w = 0; w < W; w++) c += __builtin_popcountll(row[w]);
            fsz[k] = c;
        }
    }
    std::map<int, std::vector<int>> bysize;
    for (int k = 1; k <= ns; k++) bysize[fsz[k]].push_back(k);
    std::vector<Trip> rootsplits;
    double N = 0.0;
    for (int i = 1; i <= ns; i++) {
        auto it = bysize.find(nleaves - fsz[i]);
        if (it == bysize.end()) continue;
        const uint64_t* ri = fs.data() + (size_t)i * W;
        for (int j : it->second) {
            if (j <= i) continue;
            const uint64_t* rj = fs.data() + (size_t)j * W;
            bool disjoint = true;
            for (int w = 0; w < W; w++) if (ri[w] & rj[w]) { disjoint = false; break; }
            if (!disjoint) continue;
            if (bipv[i] != bipv[j]) throw std::runtime_error(path + ": complementary clades have different counts");
            N += bipv[i];
            rootsplits.push_back(Trip{i, j, bipv[j]});
        }
    }
    if (rootsplits.empty()) throw std::runtime_error("no root splits in " + path);
    dip[G] = rootsplits;
    bipv[G] = N;
    {
        std::vector<int> all;
        const Trip& r = rootsplits.back();
        const uint64_t* ra = fs.data() + (size_t)r.a * W;
        const uint64_t* rb = fs.data() + (size_t)r.b * W;
        for (int g = 1; g <= ns; g++)
            if (((ra[g >> 6] | rb[g >> 6]) >> (g & 63)) & 1ull) all.push_back(g);
        sets[G] = all;
    }
    // CCD ctor: new ids by (size, old id)
    std::vector<int> order;
    std::vector<size_t> size_of(G + 1, 0);
    for (auto& kv : sets) { order.push_back(kv.first); size_of[kv.first] = kv.second.size(); }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const size_t sa = size_of[a], sb = size_of[b];
        return sa != sb ? sa < sb : a < b;
    });
    std::vector<int> newid(G + 1, -1);
    for (int i = 0; i < (int)order.size(); i++) newid[order[i]] = i;
    const int Gn = (int)order.size();
    F.nleaf.resize(Gn);
    F.split_off.assign(1, 0);
    {
        size_t ntrip = 0;
        for (auto& kv : dip) ntrip += kv.second.size();
        F.g1.reserve(ntrip); F.g2.reserve(ntrip); F.p.reserve(ntrip);
        F.split_off.reserve((size_t)Gn + 1);
    }
    std::vector<std::vector<uint64_t>> spmask(Gn, std::vector<uint64_t>(sp.words, 0));
    std::vector<int> species_of(G + 1, -1);  // leaf set id -> species id (looked up once per leaf, on first use)
    for (int i = 0; i < Gn; i++) {
        const int k = order[i];
        const std::vector<int>& members = sets[k];
        F.nleaf[i] = (int32_t)members.size();
        for (int g : members) {
            if (g < 0 || g > G) throw std::runtime_error("set refers to an unknown leaf in " + path);
            if (species_of[g] < 0) {
                const std::string& nm = leafname[g];
                const std::string pre = nm.substr(0, nm.find('_'));
                auto it = sp.id_of.find(pre);
                if (it == sp.id_of.end()) throw std::runtime_error("gene " + nm + " in " + path + ": species " + pre + " is not in the species tree");
                species_of[g] = it->second;
            }
            spmask[i][species_of[g] >> 6] |= 1ull << (species_of[g] & 63);
        }
        const double denom = bipv[k];
        auto dit = dip.find(k);
        if (dit != dip.end())
            for (const Trip& t : dit->second) {
                if (t.a < 1 || t.a > G || t.b < 1 || t.b > G || newid[t.a] < 0 || newid[t.b] < 0)
                    throw std::runtime_error("triple refers to an unknown clade in " + path);
                F.g1.push_back(newid[t.a]);
                F.g2.push_back(newid[t.b]);
                F.p.push_back(t.c / denom);
            }
        F.split_off.push_back((int64_t)F.g1.size());
    }
    F.compat_off.assign(1, 0);
    for (int e = 0; e < nn; e++) {
        const std::vector<uint64_t>& cm = sp.node_mask[e];
        for (int i = 0; i < Gn; i++) {
            bool sub = true;
            for (int w = 0; w < sp.words; w++) if (spmask[i][w] & ~cm[w]) { sub = false; break; }
            if (sub) F.compat.push_back(i);
        }
        F.compat_off.push_back((int64_t)F.compat.size());
    }
}

// parse all files on `nthreads` host threads (0: hardware concurrency)
static inline void parse_all(const std::vector<std::string>& paths, const Species& sp, int nn, std::vector<Family>& fams,
                             int nthreads) {
    fams.assign(paths.size(), Family());
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min<int>(nt, (int)paths.size()));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            for (size_t i = t; i < paths.size(); i += nt) {
                try { parse_family(paths[i], sp, nn, fams[i]); }
                catch (const std::exception& ex) { fams[i].error = ex.what(); }
            }
        });
    for (auto& x : th) x.join();
}

}  // namespace whale_ale
