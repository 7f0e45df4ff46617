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

static inline std::vector<std::string> split_ws(const std::string& l) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < l.size()) {
        while (i < l.size() && (l[i] == ' ' || l[i] == '\t' || l[i] == '\r')) i++;
        size_t j = i;
        while (j < l.size() && l[j] != ' ' && l[j] != '\t' && l[j] != '\r') j++;
        if (j > i) out.push_back(l.substr(i, j - i));
        i = j;
    }
    return out;
}

static inline void parse_family(const std::string& path, const Species& sp, int nn, Family& F) {
    FILE* fh = fopen(path.c_str(), "rb");
    if (!fh) throw std::runtime_error("cannot open " + path);
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fh)) > 0) text.append(buf, n);
    fclose(fh);
    std::vector<std::string> parts;
    {
        size_t a = 0;
        for (;;) {
            size_t b = text.find('#', a);
            if (b == std::string::npos) { parts.push_back(text.substr(a)); break; }
            parts.push_back(text.substr(a, b - a));
            a = b + 1;
        }
    }
    if (parts.size() != 10) throw std::runtime_error("Not a valid .ale file " + path);
    std::map<std::string, std::vector<std::string>> sec;
    for (size_t pi = 1; pi + 1 < parts.size(); pi++) {
        std::string part = parts[pi];
        for (size_t k; (k = part.find(":\t")) != std::string::npos;) part.erase(k, 2);
        std::vector<std::string> lines;
        size_t a = 0;
        while (a <= part.size()) {
            size_t b = part.find('\n', a);
            if (b == std::string::npos) b = part.size();
            std::string l = part.substr(a, b - a);
            if (!l.empty() && l.back() == '\r') l.pop_back();
            if (!l.empty()) lines.push_back(l);
            a = b + 1;
        }
        if (lines.empty()) continue;
        std::string name = lines[0];
        while (!name.empty() && (name.back() == ' ' || name.back() == '\t')) name.pop_back();
        std::replace(name.begin(), name.end(), '-', '_');
        sec[name] = std::vector<std::string>(lines.begin() + 1, lines.end());
    }
    auto need = [&](const char* k) -> const std::vector<std::string>& {
        auto it = sec.find(k);
        if (it == sec.end()) throw std::runtime_error(std::string("section ") + k + " missing in " + path);
        return it->second;
    };
    const double obs = atof(need("observations").at(0).c_str());
    std::map<int, double> bip;
    for (auto& l : need("Bip_counts")) { auto t = split_ws(l); if (t.size() >= 2) bip[atoi(t[0].c_str())] = atof(t[1].c_str()); }
    struct Trip { int a, b; double c; };
    std::map<int, std::vector<Trip>> dip;
    for (auto& l : need("Dip_counts")) {
        auto t = split_ws(l);
        if (t.size() < 4) continue;
        dip[atoi(t[0].c_str())].push_back(Trip{atoi(t[1].c_str()), atoi(t[2].c_str()), atof(t[3].c_str())});
    }
    std::map<int, std::string> leaf_name;  // leaf id -> gene name
    for (auto& l : need("leaf_id")) { auto t = split_ws(l); if (t.size() >= 2) leaf_name[atoi(t[1].c_str())] = t[0]; }
    std::map<int, std::vector<int>> sets;
    for (auto& l : need("set_id")) {
        auto t = split_ws(l);
        if (t.empty()) continue;
        std::vector<int> v;
        for (size_t i = 1; i < t.size(); i++) v.push_back(atoi(t[i].c_str()));
        sets[atoi(t[0].c_str())] = v;
    }
    // addleafclades!
    std::map<int, int> leaf2set;
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
                auto it = leaf2set.find(i);
                if (it == leaf2set.end()) throw std::runtime_error("set refers to an unknown leaf in " + path);
                i = it->second;
            }
        }
    }
    for (auto& kv : leaf2set) sets[kv.second] = std::vector<int>{kv.second};
    // addubiquitous!
    const int nleaves = (int)leafname.size();
    const int ns = (int)sets.size();
    const int G = ns + 1;
    if (G > 65535) throw std::runtime_error("more than 65535 clades (UInt16 ids) in " + path);
    const int W = (ns + 64) / 64;
    std::vector<std::vector<uint64_t>> fs(ns + 1, std::vector<uint64_t>(W, 0));
    std::vector<int> fsz(ns + 1, 0);
    for (int k = 1; k <= ns; k++) {
        auto it = sets.find(k);
        if (it == sets.end()) throw std::runtime_error("set ids are not 1..n in " + path);
        for (int g : it->second) { fs[k][g >> 6] |= 1ull << (g & 63); }
        int c = 0;
        for (uint64_t w : fs[k]) c += __builtin_popcountll(w);
        fsz[k] = c;
    }
    std::map<int, std::vector<int>> bysize;
    for (int k = 1; k <= ns; k++) bysize[fsz[k]].push_back(k);
    std::vector<Trip> rootsplits;
    double N = 0.0;
    for (int i = 1; i <= ns; i++) {
        auto it = bysize.find(nleaves - fsz[i]);
        if (it == bysize.end()) continue;
        for (int j : it->second) {
            if (j <= i) continue;
            bool disjoint = true;
            for (int w = 0; w < W; w++) if (fs[i][w] & fs[j][w]) { disjoint = false; break; }
            if (!disjoint) continue;
            if (bip[i] != bip[j]) throw std::runtime_error(path + ": complementary clades have different counts");
            N += bip[i];
            rootsplits.push_back(Trip{i, j, bip[j]});
        }
    }
    if (rootsplits.empty()) throw std::runtime_error("no root splits in " + path);
    dip[G] = rootsplits;
    bip[G] = N;
    {
        std::vector<int> all;
        const Trip& r = rootsplits.back();
        for (int g = 1; g <= ns; g++)
            if (((fs[r.a][g >> 6] | fs[r.b][g >> 6]) >> (g & 63)) & 1ull) all.push_back(g);
        sets[G] = all;
    }
    // CCD ctor: new ids by (size, old id)
    std::vector<int> order;
    for (auto& kv : sets) order.push_back(kv.first);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const size_t sa = sets[a].size(), sb = sets[b].size();
        return sa != sb ? sa < sb : a < b;
    });
    std::vector<int> newid(G + 1, -1);
    for (int i = 0; i < (int)order.size(); i++) newid[order[i]] = i;
    const int Gn = (int)order.size();
    F.nleaf.resize(Gn);
    F.split_off.assign(1, 0);
    std::vector<std::vector<uint64_t>> spmask(Gn, std::vector<uint64_t>(sp.words, 0));
    for (int i = 0; i < Gn; i++) {
        const int k = order[i];
        F.nleaf[i] = (int32_t)sets[k].size();
        for (int g : sets[k]) {
            const std::string& nm = leafname[g];
            const std::string pre = nm.substr(0, nm.find('_'));
            auto it = sp.id_of.find(pre);
            if (it == sp.id_of.end()) throw std::runtime_error("gene " + nm + " in " + path + ": species " + pre + " is not in the species tree");
            spmask[i][it->second >> 6] |= 1ull << (it->second & 63);
        }
        const double denom = bip[k];
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
