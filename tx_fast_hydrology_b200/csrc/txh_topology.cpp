// txh_topology.cpp -- topology pass + dataflow schedule builder (host, exact integers).
// See txh_topology.hpp for what each artefact replaces in the reference.
#include "txh_topology.hpp"

#include <algorithm>
#include <functional>
#include <numeric>
#include <queue>

namespace txh {

// ---------------------------------------------------------------------------
// Topology
// ---------------------------------------------------------------------------
bool Topology::build(int64_t n_, const int64_t* e, std::string& err)
{
    if (n_ <= 0 || n_ >= (int64_t(1) << 30)) { err = "n out of range"; return false; }
    n = n_;
    end.resize(n);
    for (int64_t j = 0; j < n; ++j) {
        if (e[j] < 0 || e[j] >= n) { err = "endnodes entry out of range"; return false; }
        end[j] = (int32_t)e[j];
    }
    // indegree without self-loops: muskingum.py:322-330
    indeg.assign(n, 0);
    for (int64_t j = 0; j < n; ++j)
        if (end[j] != j) indeg[end[j]] += 1;
    child_off.assign(n + 1, 0);
    for (int64_t j = 0; j < n; ++j) child_off[j + 1] = child_off[j] + indeg[j];
    child.resize(child_off[n]);
    {
        std::vector<int32_t> fill(child_off.begin(), child_off.end() - 1);
        for (int64_t j = 0; j < n; ++j)
            if (end[j] != j) child[fill[end[j]]++] = (int32_t)j;   // ascending id
    }
    heads.clear();
    for (int64_t j = 0; j < n; ++j)
        if (indeg[j] == 0) heads.push_back((int32_t)j);            // muskingum.py:444

    // Kahn by level: frontier k is exactly the set of level-k reaches.
    level.assign(n, 0);
    topo.clear(); topo.reserve(n);
    level_off.clear(); level_off.push_back(0);
    {
        std::vector<int32_t> pending(indeg), frontier(heads), next;
        int32_t lv = 0;
        while (!frontier.empty()) {
            topo.insert(topo.end(), frontier.begin(), frontier.end());
            level_off.push_back((int32_t)topo.size());
            next.clear();
            for (int32_t j : frontier) {
                int32_t d = end[j];
                if (d == j) continue;
                if (level[d] < lv + 1) level[d] = lv + 1;
                if (--pending[d] == 0) next.push_back(d);
            }
            std::sort(next.begin(), next.end());
            frontier.swap(next);
            ++lv;
        }
        nlevels = lv;
    }
    if ((int64_t)topo.size() != n) { err = "network has a cycle (not a forest of in-trees)"; return false; }

    subtree.assign(n, 1);
    for (int32_t j : topo)
        if (end[j] != j) subtree[end[j]] += subtree[j];

    main_child.assign(n, -1);
    for (int64_t j = 0; j < n; ++j) {
        int32_t best = -1;
        for (int32_t k = child_off[j]; k < child_off[j + 1]; ++k) {
            int32_t c = child[k];
            if (best < 0 || level[c] > level[best]) best = c;      // ties keep the lowest id
        }
        main_child[j] = best;
    }
    path_id.assign(n, -1); path_pos.assign(n, 0); path_len.clear();
    for (int32_t j : topo) {
        int32_t m = main_child[j];
        if (m < 0) { path_id[j] = (int32_t)path_len.size(); path_pos[j] = 0; path_len.push_back(1); }
        else {
            path_id[j] = path_id[m]; path_pos[j] = path_pos[m] + 1;
            path_len[path_id[j]] = path_pos[j] + 1;
        }
    }

    // maximal unbranched chains: v continues its upstream reach's chain iff indegree[v] == 1
    chain_id.assign(n, -1); chain_pos.assign(n, 0); chain_len.clear();
    for (int64_t j = 0; j < n; ++j)
        if (indeg[j] != 1) { chain_id[j] = (int32_t)chain_len.size(); chain_len.push_back(1); }
    for (int32_t j : topo)
        if (indeg[j] == 1) {
            int32_t u = child[child_off[j]];
            chain_id[j] = chain_id[u]; chain_pos[j] = chain_pos[u] + 1;
            chain_len[chain_id[j]] = chain_pos[j] + 1;
        }

    // the reference's visit sequence: nutils.py:72-88
    visit.clear(); visit.reserve(n);
    {
        std::vector<int32_t> work(indeg);
        for (int32_t h : heads) {
            int32_t s = h, d = end[s];
            while (work[s] == 0) {
                visit.push_back(s);
                work[d] -= 1;
                s = d; d = end[s];
                if ((int64_t)visit.size() > n) break;
            }
        }
        if ((int64_t)visit.size() != n) { err = "reference walk did not visit every reach once"; return false; }
    }
    return true;
}

// ---------------------------------------------------------------------------
// Schedule
// ---------------------------------------------------------------------------
namespace {

struct Emit {
    std::vector<int32_t> reach;       // reaches in processing order, all tasks concatenated
    std::vector<uint32_t> hdr;        // parallel to `reach`
    std::vector<uint32_t> inw;        // ROW entries hold REACH ids until positions are known
    std::vector<int32_t> in_off;      // per emitted reach
};

}  // namespace

bool Schedule::build(const Topology& t, const SchedParams& p, std::string& err)
{
    prm = p;
    const int64_t n = t.n;
    if (p.spine_cap < 1 || p.pocket_cap < 1 || p.pocket_cap > 4096 || p.long_path_min < 2 ||
        p.max_slots < 0 || p.max_slots > 30) {
        err = "bad schedule parameters"; return false;
    }
    std::vector<uint8_t> is_long(n);
    for (int64_t j = 0; j < n; ++j) is_long[j] = t.path_len[t.path_id[j]] >= p.long_path_min;

    // ---- units (tasks before ordering): members in processing order --------------------
    std::vector<int32_t> unit(n, -1);
    std::vector<int32_t> u_begin, u_len;          // into Emit arrays
    std::vector<uint8_t> u_kind;
    Emit em;
    em.reach.reserve(n); em.hdr.reserve(n); em.in_off.reserve(n + 1);

    auto push_reach = [&](int32_t j, uint32_t h, const std::vector<uint32_t>& ins) {
        em.reach.push_back(j);
        em.in_off.push_back((int32_t)em.inw.size());
        em.hdr.push_back(h | ((uint32_t)ins.size() << 6));
        em.inw.insert(em.inw.end(), ins.begin(), ins.end());
    };

    // ---- spines: long paths cut into pure-chain segments -------------------------------
    {
        const int32_t npaths = (int32_t)t.path_len.size();
        std::vector<int32_t> poff(npaths + 1, 0);
        for (int32_t q = 0; q < npaths; ++q) poff[q + 1] = poff[q] + t.path_len[q];
        std::vector<int32_t> member(n);
        for (int64_t j = 0; j < n; ++j) member[poff[t.path_id[j]] + t.path_pos[j]] = (int32_t)j;
        std::vector<uint32_t> ins;
        for (int32_t q = 0; q < npaths; ++q) {
            const int32_t len = t.path_len[q];
            if (len < p.long_path_min) continue;
            const int32_t nseg = (len + p.spine_cap - 1) / p.spine_cap;
            const int32_t seglen = (len + nseg - 1) / nseg;
            for (int32_t s0 = 0; s0 < len; s0 += seglen) {
                const int32_t s1 = std::min(len, s0 + seglen);
                const int32_t u = (int32_t)u_begin.size();
                u_begin.push_back((int32_t)em.reach.size()); u_len.push_back(s1 - s0); u_kind.push_back(0);
                for (int32_t k = s0; k < s1; ++k) {
                    const int32_t j = member[poff[q] + k];
                    unit[j] = u;
                    ins.clear();
                    uint32_t h = 0;
                    const int32_t m = t.main_child[j];
                    if (m >= 0) { if (k > s0) h |= HDR_ACC; else ins.push_back(INW_ROW | (uint32_t)m); }
                    for (int32_t c = t.child_off[j]; c < t.child_off[j + 1]; ++c)
                        if (t.child[c] != m) ins.push_back(INW_ROW | (uint32_t)t.child[c]);
                    push_reach(j, h, ins);
                }
            }
        }
    }
    n_spine = (int32_t)u_begin.size();

    // ---- pockets: side subtrees made of short paths only -------------------------------
    // split oversized pockets bottom-up; `closed[j]` marks roots of mini-trees
    std::vector<int32_t> open(n, 0), need(n, 0);
    std::vector<uint8_t> closed(n, 0);
    {
        std::vector<std::pair<int32_t, int32_t>> kids;
        for (int32_t j : t.topo) {
            if (is_long[j]) continue;
            int32_t total = 1;
            kids.clear();
            for (int32_t c = t.child_off[j]; c < t.child_off[j + 1]; ++c) {
                const int32_t ch = t.child[c];
                if (!closed[ch]) { kids.emplace_back(open[ch], ch); total += open[ch]; }
            }
            if (total > p.pocket_cap) {
                std::sort(kids.begin(), kids.end(), [](auto& a, auto& b) {
                    return a.first != b.first ? a.first > b.first : a.second < b.second; });
                for (auto& kv : kids) {
                    if (total <= p.pocket_cap) break;
                    closed[kv.second] = 1; total -= kv.first;
                }
            }
            open[j] = total;
            if (t.end[j] == j || is_long[t.end[j]]) closed[j] = 1;      // pocket root
        }
        // Sethi-Ullman scratch need over in-tree children
        std::vector<int32_t> nd;
        for (int32_t j : t.topo) {
            if (is_long[j]) continue;
            nd.clear();
            for (int32_t c = t.child_off[j]; c < t.child_off[j + 1]; ++c)
                if (!closed[t.child[c]]) nd.push_back(need[t.child[c]]);
            std::sort(nd.begin(), nd.end(), std::greater<int32_t>());
            int32_t v = 0;
            for (size_t i = 0; i < nd.size(); ++i) v = std::max(v, nd[i] + (int32_t)i);
            need[j] = v;
        }
    }
    // mini-tree roots and their bundling keys
    struct Mini { int64_t key; int32_t root; };
    std::vector<Mini> grpA, grpB, grpC;      // feeds a spine / isolated / feeds another mini-tree
    for (int64_t j = 0; j < n; ++j) {
        if (is_long[j] || !closed[j]) continue;
        const int32_t d = t.end[j];
        if (d == j) grpB.push_back({(int64_t)j, (int32_t)j});
        else if (is_long[d]) grpA.push_back({((int64_t)unit[d] << 32) | (uint32_t)t.path_pos[d], (int32_t)j});
        else grpC.push_back({(int64_t)j, (int32_t)j});
    }
    auto by_key = [](const Mini& a, const Mini& b) { return a.key != b.key ? a.key < b.key : a.root < b.root; };
    std::sort(grpA.begin(), grpA.end(), by_key);

    // DFS emission of one mini-tree into the current unit
    std::vector<int32_t> free_slots;
    int32_t slots_hi = 0;
    std::vector<int32_t> slot_of(n, -1);
    std::function<void(int32_t)> emit_tree = [&](int32_t v) {
        std::vector<std::pair<int32_t, int32_t>> kids;     // (need, child) of in-tree children
        for (int32_t c = t.child_off[v]; c < t.child_off[v + 1]; ++c)
            if (!closed[t.child[c]]) kids.emplace_back(need[t.child[c]], t.child[c]);
        std::sort(kids.begin(), kids.end(), [](auto& a, auto& b) {
            return a.first != b.first ? a.first > b.first : a.second < b.second; });
        for (size_t i = 0; i < kids.size(); ++i) {
            emit_tree(kids[i].second);
            if (i + 1 < kids.size()) {
                // park the child's outflow: patch its header with a scratch slot if one is free
                const int32_t c = kids[i].second;
                if (!free_slots.empty()) {
                    const int32_t s = free_slots.back(); free_slots.pop_back();
                    slot_of[c] = s;
                    slots_hi = std::max(slots_hi, s + 1);
                    em.hdr.back() |= (uint32_t)(s + 1) << 1;   // c is the reach just emitted
                } else {
                    row_fallbacks += 1;
                }
            }
        }
        std::vector<uint32_t> ins;
        uint32_t h = 0;
        if (!kids.empty()) h |= HDR_ACC;                       // last in-tree child was emitted just before v
        for (size_t i = 0; i + 1 < kids.size(); ++i) {
            const int32_t c = kids[i].second;
            if (slot_of[c] >= 0) { ins.push_back((uint32_t)slot_of[c]); free_slots.push_back(slot_of[c]); slot_of[c] = -1; }
            else ins.push_back(INW_ROW | (uint32_t)c);
        }
        for (int32_t c = t.child_off[v]; c < t.child_off[v + 1]; ++c)
            if (closed[t.child[c]]) ins.push_back(INW_ROW | (uint32_t)t.child[c]);
        unit[v] = (int32_t)u_begin.size() - 1;
        push_reach(v, h, ins);
    };
    auto reset_slots = [&]() {
        free_slots.clear();
        for (int32_t s = p.max_slots - 1; s >= 0; --s) free_slots.push_back(s);
    };
    auto pack = [&](const std::vector<Mini>& grp, bool bundle, bool same_high_key) {
        int32_t cur = -1, cur_size = 0; int64_t cur_hi = -1;
        for (const Mini& m : grp) {
            const int32_t sz = open[m.root];
            const int64_t hi = m.key >> 32;
            const bool fits = bundle && cur >= 0 && cur_size + sz <= p.pocket_cap &&
                              (!same_high_key || hi == cur_hi);
            if (!fits) {
                cur = (int32_t)u_begin.size();
                u_begin.push_back((int32_t)em.reach.size()); u_len.push_back(0); u_kind.push_back(1);
                cur_size = 0; cur_hi = hi;
            }
            reset_slots();
            emit_tree(m.root);
            cur_size += sz;
            u_len[cur] = (int32_t)em.reach.size() - u_begin[cur];
        }
    };
    pack(grpC, false, false);
    pack(grpA, true, true);
    pack(grpB, true, false);
    n_pocket = (int32_t)u_begin.size() - n_spine;
    slots_used = slots_hi;
    if ((int64_t)em.reach.size() != n) { err = "internal: schedule does not cover every reach"; return false; }
    em.in_off.push_back((int32_t)em.inw.size());

    // ---- task DAG -----------------------------------------------------------------------
    const int32_t nu = (int32_t)u_begin.size();
    std::vector<std::vector<int32_t>> prod(nu), cons(nu);
    for (int32_t u = 0; u < nu; ++u) {
        std::vector<int32_t>& pr = prod[u];
        for (int32_t e = u_begin[u]; e < u_begin[u] + u_len[u]; ++e)
            for (int32_t w = em.in_off[e]; w < em.in_off[e + 1]; ++w)
                if (em.inw[w] & INW_ROW) {
                    const int32_t pu = unit[em.inw[w] & ~INW_ROW];
                    if (pu != u) pr.push_back(pu);
                }
        std::sort(pr.begin(), pr.end());
        pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
        for (int32_t pu : pr) cons[pu].push_back(u);
    }
    // plain Kahn for a first topological order, then critical-path-to-sink costs
    std::vector<int32_t> order0; order0.reserve(nu);
    {
        std::vector<int32_t> pend(nu);
        for (int32_t u = 0; u < nu; ++u) { pend[u] = (int32_t)prod[u].size(); if (!pend[u]) order0.push_back(u); }
        for (size_t k = 0; k < order0.size(); ++k)
            for (int32_t c : cons[order0[k]]) if (--pend[c] == 0) order0.push_back(c);
        if ((int32_t)order0.size() != nu) { err = "internal: task graph has a cycle"; return false; }
    }
    std::vector<int64_t> cp(nu, 0);
    std::vector<int32_t> cpt(nu, 0);
    for (int32_t k = nu - 1; k >= 0; --k) {
        const int32_t u = order0[k];
        int64_t best = 0; int32_t bt = 0;
        for (int32_t c : cons[u]) { if (cp[c] > best) best = cp[c]; if (cpt[c] > bt) bt = cpt[c]; }
        cp[u] = best + 24 + u_len[u];
        cpt[u] = bt + 1;
    }
    cp_cost = 0; cp_tasks = 0;
    for (int32_t u = 0; u < nu; ++u) { cp_cost = std::max(cp_cost, cp[u]); cp_tasks = std::max(cp_tasks, cpt[u]); }
    // claim order: topological, longest remaining critical path first
    std::vector<int32_t> order; order.reserve(nu);
    {
        auto cmp = [&](int32_t a, int32_t b) { return cp[a] != cp[b] ? cp[a] < cp[b] : a > b; };
        std::priority_queue<int32_t, std::vector<int32_t>, decltype(cmp)> pq(cmp);
        std::vector<int32_t> pend(nu);
        for (int32_t u = 0; u < nu; ++u) { pend[u] = (int32_t)prod[u].size(); if (!pend[u]) pq.push(u); }
        while (!pq.empty()) {
            const int32_t u = pq.top(); pq.pop();
            order.push_back(u);
            for (int32_t c : cons[u]) if (--pend[c] == 0) pq.push(c);
        }
    }
    std::vector<int32_t> rank(nu);
    for (int32_t k = 0; k < nu; ++k) rank[order[k]] = k;

    // ---- positions and device descriptors -------------------------------------------------
    pos_of_reach.assign(n, -1); reach_of_pos.assign(n, -1);
    task_of_pos.assign(n, -1);
    tasks.assign(nu, TaskDesc{});
    task_kind.assign(nu, 0);
    hdr.assign(n, 0);
    inw.clear(); inw.reserve(em.inw.size());
    deps.clear();
    {
        int32_t pos = 0;
        for (int32_t k = 0; k < nu; ++k) {
            const int32_t u = order[k];
            tasks[k].begin = pos; tasks[k].len = u_len[u];
            task_kind[k] = u_kind[u];
            for (int32_t e = u_begin[u]; e < u_begin[u] + u_len[u]; ++e, ++pos) {
                pos_of_reach[em.reach[e]] = pos; reach_of_pos[pos] = em.reach[e]; task_of_pos[pos] = k;
            }
        }
        for (int32_t k = 0; k < nu; ++k) {
            const int32_t u = order[k];
            tasks[k].in_off = (int32_t)inw.size();
            int32_t pos = tasks[k].begin;
            for (int32_t e = u_begin[u]; e < u_begin[u] + u_len[u]; ++e, ++pos) {
                hdr[pos] = em.hdr[e];
                for (int32_t w = em.in_off[e]; w < em.in_off[e + 1]; ++w) {
                    uint32_t x = em.inw[w];
                    if (x & INW_ROW) x = INW_ROW | (uint32_t)pos_of_reach[x & ~INW_ROW];
                    inw.push_back(x);
                }
            }
            tasks[k].dep_off = (int32_t)deps.size();
            tasks[k].n_raw = (int32_t)prod[u].size();
            tasks[k].n_war = (int32_t)cons[u].size();
            for (int32_t pu : prod[u]) {
                if (rank[pu] >= k) { err = "internal: producer ordered after consumer"; return false; }
                deps.push_back(rank[pu]);
            }
            for (int32_t cu : cons[u]) deps.push_back(rank[cu]);
        }
    }

    // ---- position-space CSR of upstream rows + level lists -------------------------------
    up_off.assign(n + 1, 0); up_pos.resize(t.child.size());
    is_outlet_pos.assign(n, 0);
    for (int64_t q = 0; q < n; ++q) {
        const int32_t j = reach_of_pos[q];
        up_off[q + 1] = up_off[q] + (t.child_off[j + 1] - t.child_off[j]);
        int32_t w = up_off[q];
        for (int32_t c = t.child_off[j]; c < t.child_off[j + 1]; ++c) up_pos[w++] = pos_of_reach[t.child[c]];
        is_outlet_pos[q] = t.end[j] == j;
    }
    lvl_off.assign(t.level_off.begin(), t.level_off.end());
    lvl_pos.resize(n);
    for (int32_t l = 0; l < t.nlevels; ++l) {
        for (int32_t k = t.level_off[l]; k < t.level_off[l + 1]; ++k) lvl_pos[k] = pos_of_reach[t.topo[k]];
        std::sort(lvl_pos.begin() + t.level_off[l], lvl_pos.begin() + t.level_off[l + 1]);
    }
    return true;
}

}  // namespace txh
