// Candidate enumeration (host, C++): splice graph in the reference's DOT mini-format -> every source-to-sink path as a
// candidate transcript.  SURVEY.md 8f row f2; replaces, behind the C ABI of include/bayesembler_b200.h,
//   AllPaths::AllPaths / parseGraphviz / parseVertexProperties      allPaths.cpp:42-113
//   AllPaths::findFullPaths / findAllPathsSource                     allPaths.cpp:116-343
// and produces the arrays bayes_raw_locus takes (the likelihood builder, row f1).
//
// Same candidates in the same order as the reference.  What changes is how they are found:
//   * the reference parses the text with Boost.Graph's read_graphviz into an adjacency_list; here a small scanner for
//     the three statement kinds the producer writes (assembler.cpp:764-861) fills flat arrays.  Vertex order = order of
//     the node ids as strings (read_graphviz collects nodes in a std::map before adding them), out-edge order = order of
//     the edge statements;
//   * the reference finds out that a graph has more than --max-candidate-number paths by enumerating them, raises the
//     junction threshold by ONE, deletes the junctions below it and starts over (allPaths.cpp:173-289) -- on the
//     config-3 class of graphs that is thousands of aborted enumerations.  Here the number of paths under a threshold is
//     a saturating dynamic programme over the DAG (one pass over the edges), thresholds at which no junction disappears
//     are skipped, and the paths are enumerated exactly once, at the threshold the reference ends on;
//   * the recursive enumeration that copies every partial path list at every level (:307-343) is an iterative
//     depth-first walk with one shared path stack; paths come out in the same (source order, adjacency order) sequence.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "internal.h"

struct bayes_path_set {
    std::string graph_name, reference, strand;
    std::vector<int32_t> exon_ptr, exon_start, exon_end, length, is_premrna;
    // vertices, sources, sinks (as logged at allPaths.cpp:132-134, i.e. of the graph as read), junction-threshold
    // adjustments (:317), paths (:301), final junction threshold
    int32_t stats[6] = {0, 0, 0, 0, 0, 1};
};

namespace {

struct Node {
    std::string id;
    std::string label;
    bool has_label = false;
};
struct Edge {
    int src, dst, coverage;
};

struct Graph {
    std::string name;
    std::vector<Node> nodes;  // sorted by id
    std::vector<int32_t> start, end;
    std::vector<Edge> edges;  // statement order
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }
inline bool is_id_char(char c) {
    return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_' || c == '.';
}

// The text the producer writes (assembler.cpp:764, 824, 849-858):
//   digraph <name> {\n   <id> [label="ref;strand;start;end"]\n ...   <id>-><id> [coverage="n"]\n ...   }\n
struct Scanner {
    const char *p, *e;
    std::string err;
    void skip() { while (p < e && (is_space(*p) || *p == ';')) p++; }
    bool ident(std::string &out) {
        out.clear();
        if (p < e && *p == '"') {
            for (p++; p < e && *p != '"'; p++) {
                if (*p == '\\' && p + 1 < e) p++;
                out += *p;
            }
            if (p >= e) return false;
            p++;
            return true;
        }
        while (p < e && (is_id_char(*p) || (*p == '-' && !(p + 1 < e && p[1] == '>')))) out += *p++;
        return !out.empty();
    }
};

struct Stmt {
    std::string a, b, label;
    bool has_label = false, has_cov = false;
    int coverage = 0;
};

int parse_dot(const char *text, Graph &g, std::string &err) {
    Scanner s{text, text + strlen(text), std::string()};
    // graph name: the reference takes it from a line of the form "digraph <name> {" (allPaths.cpp:69-74)
    s.skip();
    if ((size_t)(s.e - s.p) < 8 || strncmp(s.p, "digraph", 7) != 0 || !is_space(s.p[7])) { err = "DOT text does not start with 'digraph <name> {'"; return BAYES_E_INVALID; }
    s.p += 7;
    s.skip();
    const char *n0 = s.p;
    while (s.p < s.e && !is_space(*s.p) && *s.p != '{') s.p++;
    g.name.assign(n0, s.p);
    s.skip();
    if (g.name.empty() || s.p >= s.e || *s.p != '{') { err = "DOT text does not start with 'digraph <name> {'"; return BAYES_E_INVALID; }
    s.p++;
    std::vector<Stmt> stmts;
    for (;;) {
        s.skip();
        if (s.p >= s.e) { err = "DOT text ends before '}'"; return BAYES_E_INVALID; }
        if (*s.p == '}') break;
        Stmt st;
        if (!s.ident(st.a)) { err = "DOT: node id expected"; return BAYES_E_INVALID; }
        s.skip();
        if (s.p + 1 < s.e && s.p[0] == '-' && s.p[1] == '>') {
            s.p += 2;
            s.skip();
            if (!s.ident(st.b)) { err = "DOT: node id expected after '->'"; return BAYES_E_INVALID; }
            s.skip();
        }
        if (s.p < s.e && *s.p == '[') {
            s.p++;
            for (;;) {
                while (s.p < s.e && (is_space(*s.p) || *s.p == ',')) s.p++;
                if (s.p < s.e && *s.p == ']') { s.p++; break; }
                std::string k, v;
                if (!s.ident(k)) { err = "DOT: attribute name expected"; return BAYES_E_INVALID; }
                while (s.p < s.e && is_space(*s.p)) s.p++;
                if (s.p >= s.e || *s.p != '=') { err = "DOT: '=' expected"; return BAYES_E_INVALID; }
                s.p++;
                while (s.p < s.e && is_space(*s.p)) s.p++;
                if (!s.ident(v)) { err = "DOT: attribute value expected"; return BAYES_E_INVALID; }
                if (k == "label") { st.label = v; st.has_label = true; }
                else if (k == "coverage") { st.coverage = atoi(v.c_str()); st.has_cov = true; }
            }
        }
        stmts.push_back(st);
    }
    // nodes in the order of their ids as strings; a node statement sets the label, an edge statement only mentions ids
    std::vector<std::string> ids;
    ids.reserve(stmts.size() * 2);
    for (const Stmt &st : stmts) {
        ids.push_back(st.a);
        if (!st.b.empty()) ids.push_back(st.b);
    }
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    g.nodes.resize(ids.size());
    for (size_t i = 0; i < ids.size(); i++) g.nodes[i].id = ids[i];
    auto find = [&](const std::string &id) { return (int)(std::lower_bound(ids.begin(), ids.end(), id) - ids.begin()); };
    for (const Stmt &st : stmts) {
        if (st.b.empty()) {
            if (st.has_label) {
                Node &nd = g.nodes[find(st.a)];
                nd.label = st.label;
                nd.has_label = true;
            }
        } else {
            g.edges.push_back(Edge{find(st.a), find(st.b), st.coverage});
        }
    }
    return BAYES_OK;
}

// "ref;strand;start;end" (allPaths.cpp:88-108); one reference and one strand per graph (:110-118)
int parse_labels(Graph &g, std::string &reference, std::string &strand, std::string &err) {
    const int V = (int)g.nodes.size();
    g.start.resize(V);
    g.end.resize(V);
    for (int v = 0; v < V; v++) {
        const std::string &l = g.nodes[v].label;
        std::vector<std::string> f(1);
        for (char c : l) {
            if (c == ';') f.push_back(std::string());
            else f.back() += c;
        }
        if (!g.nodes[v].has_label || f.size() < 4) { err = "DOT: vertex '" + g.nodes[v].id + "' has no label of the form ref;strand;start;end"; return BAYES_E_INVALID; }
        g.start[v] = atoi(f[2].c_str());
        g.end[v] = atoi(f[3].c_str());
        if (v == 0) { reference = f[0]; strand = f[1]; }
        else if (f[0] != reference || f[1] != strand) { err = "DOT: vertices of one graph lie on different references / strands"; return BAYES_E_INVALID; }
        if (g.end[v] < g.start[v]) { err = "DOT: vertex '" + g.nodes[v].id + "' ends before it starts"; return BAYES_E_INVALID; }
    }
    return BAYES_OK;
}

// Out-edges that survive the junction threshold, per vertex in statement order (CSR)
struct Active {
    std::vector<int> ptr, dst;
    std::vector<char> has_in;
    int n_edges = 0, min_cov = 0;
};

void active_edges(const Graph &g, int V, const std::vector<char> &alive, int threshold, Active &a) {
    a.ptr.assign(V + 1, 0);
    a.has_in.assign(V, 0);
    a.n_edges = 0;
    a.min_cov = 0;
    for (const Edge &e : g.edges) {
        if (!alive[e.src] || !alive[e.dst] || threshold > e.coverage) continue;  // :185-188
        a.ptr[e.src + 1]++;
        a.has_in[e.dst] = 1;
        if (a.n_edges == 0 || e.coverage < a.min_cov) a.min_cov = e.coverage;
        a.n_edges++;
    }
    for (int v = 0; v < V; v++) a.ptr[v + 1] += a.ptr[v];
    a.dst.resize(a.n_edges);
    std::vector<int> fill(a.ptr.begin(), a.ptr.end() - 1);
    for (const Edge &e : g.edges) {
        if (!alive[e.src] || !alive[e.dst] || threshold > e.coverage) continue;
        a.dst[fill[e.src]++] = e.dst;
    }
}

// Number of paths from every vertex to a sink, saturating at cap; false when the graph has a cycle (the reference would
// not terminate on one; its producer only writes edges from a lower to a higher exon number).
bool count_paths(const Active &a, int V, int64_t cap, std::vector<int64_t> &cnt) {
    cnt.assign(V, -1);
    std::vector<char> on_stack(V, 0);
    std::vector<int> stack, pos;
    for (int root = 0; root < V; root++) {
        if (cnt[root] >= 0) continue;
        stack.assign(1, root);
        pos.assign(1, a.ptr[root]);
        on_stack[root] = 1;
        while (!stack.empty()) {
            const int v = stack.back();
            int &k = pos.back();
            if (k < a.ptr[v + 1]) {
                const int w = a.dst[k++];
                if (cnt[w] >= 0) continue;
                if (on_stack[w]) return false;
                on_stack[w] = 1;
                stack.push_back(w);
                pos.push_back(a.ptr[w]);
            } else {
                int64_t c = 0;
                if (a.ptr[v] == a.ptr[v + 1]) c = 1;
                for (int j = a.ptr[v]; j < a.ptr[v + 1]; j++) c = std::min(cap, c + cnt[a.dst[j]]);
                cnt[v] = c;
                on_stack[v] = 0;
                stack.pop_back();
                pos.pop_back();
            }
        }
    }
    return true;
}

}  // namespace

extern "C" {

int bayes_all_paths_build(const char *dot, int32_t no_pre_mrna, int32_t max_candidate_number, bayes_path_set **out) {
    if (!dot || !out || max_candidate_number < 0) {
        bayes::set_error("bayes_all_paths_build: null argument or negative max_candidate_number");
        return BAYES_E_INVALID;
    }
    *out = nullptr;
    Graph g;
    std::string err;
    bayes_path_set *ps = new (std::nothrow) bayes_path_set;
    if (!ps) { bayes::set_error("bayes_all_paths_build: out of memory"); return BAYES_E_NOMEM; }
    int rc = parse_dot(dot, g, err);
    if (rc == BAYES_OK) rc = parse_labels(g, ps->reference, ps->strand, err);
    if (rc != BAYES_OK) {
        bayes::set_error("bayes_all_paths_build: " + err);
        delete ps;
        return rc;
    }
    ps->graph_name = g.name;
    const int V = (int)g.nodes.size();
    std::vector<char> alive(V, 1);

    // the log block of allPaths.cpp:123-134: sources / sinks of the graph as read, before any junction is removed
    {
        std::vector<char> non_source(V, 0), non_sink(V, 0);
        for (const Edge &e : g.edges) { non_source[e.dst] = 1; non_sink[e.src] = 1; }
        int ns = 0, nk = 0;
        for (int v = 0; v < V; v++) { ns += non_source[v]; nk += non_sink[v]; }
        ps->stats[0] = V;
        ps->stats[1] = V - ns;
        ps->stats[2] = V - nk;
        // :136-167: a graph that is one exon plus an identical pre-mRNA keeps only the exon
        if (ns == 0 && nk == 0 && !no_pre_mrna && V == 2) {
            int v0 = -1, vp = -1;
            for (int v = 0; v < V; v++) {
                if (g.nodes[v].id == "0") v0 = v;
                if (g.nodes[v].id == "pre_mrna") vp = v;
            }
            if (v0 >= 0 && vp >= 0 && g.start[v0] == g.start[vp] && g.end[v0] == g.end[vp]) alive[vp] = 0;
        }
    }

    const int64_t max_c = max_candidate_number;
    int threshold = 1;
    Active act;
    std::vector<int64_t> cnt;
    int last_source = V;  // enumerate the sources below this vertex index
    for (;;) {
        active_edges(g, V, alive, threshold, act);
        if (!count_paths(act, V, max_c + 1, cnt)) {
            bayes::set_error("bayes_all_paths_build: the splice graph has a cycle");
            delete ps;
            return BAYES_E_INVALID;
        }
        // :198-217: sources in vertex order, the running path count decides
        int64_t total = 0;
        int stop_at = V;
        for (int v = 0; v < V; v++) {
            if (!alive[v] || act.has_in[v]) continue;
            total = std::min(max_c + 1, total + cnt[v]);
            if (total > max_c) { stop_at = v; break; }
        }
        if (total <= max_c) { last_source = V; break; }
        if (act.n_edges == 0) {
            // :291-294: nothing left to remove; the reference keeps the candidates of the sources it had finished
            threshold++;
            last_source = stop_at;
            break;
        }
        threshold = act.min_cov + 1;  // the next threshold at which a junction disappears (the reference steps by one)
    }
    ps->stats[3] = threshold - 1;
    ps->stats[5] = threshold;

    // enumerate: sources in vertex order, depth first in adjacency order (:307-343), exons merged when adjacent (:237-244)
    ps->exon_ptr.push_back(0);
    std::vector<int> stack, pos;
    for (int src = 0; src < last_source; src++) {
        if (!alive[src] || act.has_in[src]) continue;
        stack.assign(1, src);
        pos.assign(1, act.ptr[src]);
        while (!stack.empty()) {
            const int v = stack.back();
            if (act.ptr[v] == act.ptr[v + 1]) {  // sink: one path
                const bool pre = g.nodes[stack[0]].id == "pre_mrna";
                if (pre && stack.size() != 1) {
                    bayes::set_error("bayes_all_paths_build: 'pre_mrna' must be an isolated vertex");
                    delete ps;
                    return BAYES_E_INVALID;
                }
                if (!(pre && no_pre_mrna)) {
                    int len = 0;
                    for (size_t j = 0; j < stack.size(); j++) {
                        const int u = stack[j];
                        len += g.end[u] - g.start[u] + 1;
                        if (j > 0 && g.end[stack[j - 1]] + 1 == g.start[u]) ps->exon_end.back() = g.end[u];
                        else { ps->exon_start.push_back(g.start[u]); ps->exon_end.push_back(g.end[u]); }
                    }
                    ps->exon_ptr.push_back((int32_t)ps->exon_start.size());
                    ps->length.push_back(len);
                    ps->is_premrna.push_back(pre ? 1 : 0);
                }
                stack.pop_back();
                pos.pop_back();
                continue;
            }
            int &k = pos.back();
            if (k < act.ptr[v + 1]) {
                const int w = act.dst[k++];
                stack.push_back(w);
                pos.push_back(act.ptr[w]);
            } else {
                stack.pop_back();
                pos.pop_back();
            }
        }
    }
    ps->stats[4] = (int32_t)ps->length.size();
    *out = ps;
    return BAYES_OK;
}

void bayes_path_set_destroy(bayes_path_set *ps) { delete ps; }

int32_t bayes_path_set_num_candidates(const bayes_path_set *ps) { return ps ? (int32_t)ps->length.size() : 0; }
int32_t bayes_path_set_num_exons(const bayes_path_set *ps) { return ps ? (int32_t)ps->exon_start.size() : 0; }
const char *bayes_path_set_graph_name(const bayes_path_set *ps) { return ps ? ps->graph_name.c_str() : ""; }
const char *bayes_path_set_reference(const bayes_path_set *ps) { return ps ? ps->reference.c_str() : ""; }
const char *bayes_path_set_strand(const bayes_path_set *ps) { return ps ? ps->strand.c_str() : ""; }

int bayes_path_set_export(const bayes_path_set *ps, int32_t *exon_ptr, int32_t *exon_start, int32_t *exon_end, int32_t *cand_length,
                          int32_t *is_premrna) {
    if (!ps) { bayes::set_error("bayes_path_set_export: null path set"); return BAYES_E_INVALID; }
    if (exon_ptr) std::copy(ps->exon_ptr.begin(), ps->exon_ptr.end(), exon_ptr);
    if (exon_start) std::copy(ps->exon_start.begin(), ps->exon_start.end(), exon_start);
    if (exon_end) std::copy(ps->exon_end.begin(), ps->exon_end.end(), exon_end);
    if (cand_length) std::copy(ps->length.begin(), ps->length.end(), cand_length);
    if (is_premrna) std::copy(ps->is_premrna.begin(), ps->is_premrna.end(), is_premrna);
    return BAYES_OK;
}

int bayes_path_set_stats(const bayes_path_set *ps, int32_t *stats6) {
    if (!ps || !stats6) { bayes::set_error("bayes_path_set_stats: null argument"); return BAYES_E_INVALID; }
    for (int i = 0; i < 6; i++) stats6[i] = ps->stats[i];
    return BAYES_OK;
}

// Candidate::name (allPaths.cpp:254, 270): "<graph>_pre" for the pre-mRNA, "<graph>_seq<idx>" otherwise
int bayes_path_set_candidate_name(const bayes_path_set *ps, int32_t i, char *buf, int32_t cap) {
    if (!ps || !buf || cap <= 0 || i < 0 || i >= (int32_t)ps->length.size()) {
        bayes::set_error("bayes_path_set_candidate_name: bad argument");
        return BAYES_E_INVALID;
    }
    std::string nm = ps->graph_name + (ps->is_premrna[i] ? std::string("_pre") : "_seq" + std::to_string(i));
    if ((int32_t)nm.size() + 1 > cap) { bayes::set_error("bayes_path_set_candidate_name: buffer too small"); return BAYES_E_INVALID; }
    memcpy(buf, nm.c_str(), nm.size() + 1);
    return (int)nm.size();
}

}  // extern "C"
