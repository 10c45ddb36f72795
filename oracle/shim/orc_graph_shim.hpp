/*
 * oracle/shim/orc_graph_shim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-in for the subset of Boost.Graph / Boost.Regex / Boost.StringAlgo that the reference's allPaths.{h,cpp} use, so
 * that this UNMODIFIED translation unit compiles where it lies (oracle/Makefile -> oracle/_ref/libbayesref_f2.so) and can
 * check the product's candidate enumeration.  Boost is absent from this image and not pinned by
 * the reference.  The behaviour that matters for candidate ORDER is taken from Boost.Graph's published implementation:
 *   - read_graphviz (the non-Spirit parser, the only one with a std::string overload): every node of the DOT text is
 *     collected in a std::map keyed by its id, vertices are then added in that map's order (lexicographic by id string:
 *     "0" "1" "10" "11" "2" ... "pre_mrna"), edges afterwards in the order of the text
 *     (boost/graph/detail/read_graphviz_new.hpp, translate_results_to_graph);
 *   - adjacency_list<vecS, vecS, directedS>: vertex descriptors are indices; out-edges of a vertex keep insertion order;
 *     edges(g) walks the vertices in index order and each one's out-edges; remove_edge(e, g) erases that one out-edge
 *     and keeps the order of the rest; remove_vertex(v, g) shifts every larger index down by one.
 * "parity unpinned" against a Boost-linked binary, like every Boost-dependent step of this repository.
 */
#ifndef ORC_GRAPH_SHIM_HPP
#define ORC_GRAPH_SHIM_HPP

#include <cassert>
#include <cstdlib>
#include <map>
#include <regex>
#include <string>
#include <utility>
#include <vector>

namespace boost {

struct vecS {};
struct directedS {};
struct no_property {};
enum vertex_name_t { vertex_name };
enum vertex_color_t { vertex_color };
enum edge_weight_t { edge_weight };
template <class Tag, class T, class Next = no_property> struct property {};

namespace orc_graph {
struct OutEdge {
    std::size_t target;
    std::size_t id;  // index into edge_weight (never reused)
};
struct VertexRec {
    std::string name, color;
    std::vector<OutEdge> out;
};
struct EdgeDesc {
    std::size_t src, tgt, id;
    bool operator==(const EdgeDesc &o) const { return id == o.id; }
    bool operator!=(const EdgeDesc &o) const { return id != o.id; }
};
}  // namespace orc_graph

/* adjacency_list<vecS, vecS, directedS, {name, color : string}, {weight : int}> -- the one instantiation allPaths.h makes */
template <class OutEdgeList, class VertexList, class Directed, class VertexProp, class EdgeProp> class adjacency_list {
  public:
    std::vector<orc_graph::VertexRec> v;
    std::vector<int> edge_weight_store;

    struct vertex_iterator {
        std::size_t i;
        std::size_t operator*() const { return i; }
        vertex_iterator &operator++() { ++i; return *this; }
        vertex_iterator operator++(int) { vertex_iterator t = *this; ++i; return t; }
        bool operator!=(const vertex_iterator &o) const { return i != o.i; }
        bool operator==(const vertex_iterator &o) const { return i == o.i; }
    };
    struct edge_iterator {
        const adjacency_list *g;
        std::size_t vi, ei;
        void settle() { while (vi < g->v.size() && ei >= g->v[vi].out.size()) { ++vi; ei = 0; } }
        orc_graph::EdgeDesc operator*() const { const orc_graph::OutEdge &e = g->v[vi].out[ei]; return orc_graph::EdgeDesc{vi, e.target, e.id}; }
        edge_iterator &operator++() { ++ei; settle(); return *this; }
        edge_iterator operator++(int) { edge_iterator t = *this; ++*this; return t; }
        bool operator!=(const edge_iterator &o) const { return vi != o.vi || ei != o.ei; }
        bool operator==(const edge_iterator &o) const { return !(*this != o); }
    };
    struct adjacency_iterator {
        const adjacency_list *g;
        std::size_t vi, ei;
        std::size_t operator*() const { return g->v[vi].out[ei].target; }
        adjacency_iterator &operator++() { ++ei; return *this; }
        adjacency_iterator operator++(int) { adjacency_iterator t = *this; ++ei; return t; }
        bool operator!=(const adjacency_iterator &o) const { return ei != o.ei; }
        bool operator==(const adjacency_iterator &o) const { return ei == o.ei; }
    };
};

template <class G> struct graph_traits {
    typedef std::size_t vertex_descriptor;
    typedef orc_graph::EdgeDesc edge_descriptor;
    typedef typename G::vertex_iterator vertex_iterator;
    typedef typename G::edge_iterator edge_iterator;
    typedef typename G::adjacency_iterator adjacency_iterator;
    typedef std::size_t vertices_size_type;
};

/* property maps: a pointer to the graph + which field */
template <class G, class Tag> struct property_map;
template <class G> struct property_map<G, vertex_name_t> {
    struct type {
        G *g = nullptr;
        std::string &operator[](std::size_t vd) const { return g->v[vd].name; }
    };
};
template <class G> struct property_map<G, vertex_color_t> {
    struct type {
        G *g = nullptr;
        std::string &operator[](std::size_t vd) const { return g->v[vd].color; }
    };
};
template <class G> struct property_map<G, edge_weight_t> {
    struct type {
        G *g = nullptr;
        int &operator[](const orc_graph::EdgeDesc &e) const { return g->edge_weight_store[e.id]; }
    };
};
#define ORC_G adjacency_list<A, B, C, D, E>
#define ORC_GT template <class A, class B, class C, class D, class E>
ORC_GT typename property_map<ORC_G, vertex_name_t>::type get(vertex_name_t, ORC_G &g) { typename property_map<ORC_G, vertex_name_t>::type m; m.g = &g; return m; }
ORC_GT typename property_map<ORC_G, vertex_color_t>::type get(vertex_color_t, ORC_G &g) { typename property_map<ORC_G, vertex_color_t>::type m; m.g = &g; return m; }
ORC_GT typename property_map<ORC_G, edge_weight_t>::type get(edge_weight_t, ORC_G &g) { typename property_map<ORC_G, edge_weight_t>::type m; m.g = &g; return m; }

ORC_GT std::pair<typename ORC_G::vertex_iterator, typename ORC_G::vertex_iterator> vertices(const ORC_G &g) {
    return std::make_pair(typename ORC_G::vertex_iterator{0}, typename ORC_G::vertex_iterator{g.v.size()});
}
ORC_GT std::pair<typename ORC_G::edge_iterator, typename ORC_G::edge_iterator> edges(const ORC_G &g) {
    typename ORC_G::edge_iterator b{&g, 0, 0}, e{&g, g.v.size(), 0};
    b.settle();
    return std::make_pair(b, e);
}
ORC_GT std::pair<typename ORC_G::adjacency_iterator, typename ORC_G::adjacency_iterator> adjacent_vertices(std::size_t s, const ORC_G &g) {
    return std::make_pair(typename ORC_G::adjacency_iterator{&g, s, 0}, typename ORC_G::adjacency_iterator{&g, s, g.v[s].out.size()});
}
ORC_GT std::size_t source(const orc_graph::EdgeDesc &e, const ORC_G &) { return e.src; }
ORC_GT std::size_t target(const orc_graph::EdgeDesc &e, const ORC_G &) { return e.tgt; }
ORC_GT std::size_t num_vertices(const ORC_G &g) { return g.v.size(); }
ORC_GT std::size_t num_edges(const ORC_G &g) {
    std::size_t n = 0;
    for (std::size_t i = 0; i < g.v.size(); i++) n += g.v[i].out.size();
    return n;
}
ORC_GT void remove_edge(const orc_graph::EdgeDesc &e, ORC_G &g) {
    std::vector<orc_graph::OutEdge> &o = g.v[e.src].out;
    for (std::size_t i = 0; i < o.size(); i++)
        if (o[i].id == e.id) { o.erase(o.begin() + i); return; }
}
ORC_GT void remove_vertex(std::size_t vd, ORC_G &g) {
    assert(g.v[vd].out.empty());
    g.v.erase(g.v.begin() + vd);
    for (std::size_t i = 0; i < g.v.size(); i++)
        for (std::size_t j = 0; j < g.v[i].out.size(); j++) {
            assert(g.v[i].out[j].target != vd);
            if (g.v[i].out[j].target > vd) g.v[i].out[j].target--;
        }
}

/* dynamic_properties: the attribute -> field binding is fixed in read_graphviz below (label -> vertex_color,
 * coverage -> edge_weight, node id -> vertex_name), which is the binding allPaths.cpp:56-66 sets up */
struct dynamic_properties {
    template <class M> dynamic_properties &property(const std::string &, M) { return *this; }
};

namespace orc_graph {
/* DOT subset: `digraph NAME {`, then statements `ID [k="v", ...]` / `ID -> ID [k="v"]` separated by newlines or ';', `}` */
struct DotStmt {
    std::string a, b;  // b empty: node statement
    std::vector<std::pair<std::string, std::string> > attrs;
};
inline bool dot_parse(const std::string &s, std::vector<DotStmt> &out) {
    std::size_t i = s.find('{');
    if (i == std::string::npos) return false;
    i++;
    const std::size_t n = s.size();
    auto ws = [&]() { while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r' || s[i] == ';')) i++; };
    auto id = [&](std::string &o) {
        o.clear();
        if (i < n && s[i] == '"') {
            i++;
            while (i < n && s[i] != '"') { if (s[i] == '\\' && i + 1 < n) i++; o += s[i++]; }
            if (i >= n) return false;
            i++;
            return true;
        }
        while (i < n && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '.' || (s[i] == '-' && !(i + 1 < n && s[i + 1] == '>')))) o += s[i++];
        return !o.empty();
    };
    for (;;) {
        ws();
        if (i >= n) return false;
        if (s[i] == '}') return true;
        DotStmt st;
        if (!id(st.a)) return false;
        ws();
        if (i + 1 < n && s[i] == '-' && s[i + 1] == '>') {
            i += 2;
            ws();
            if (!id(st.b)) return false;
            ws();
        }
        if (i < n && s[i] == '[') {
            i++;
            for (;;) {
                while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == ',' || s[i] == '\n')) i++;
                if (i < n && s[i] == ']') { i++; break; }
                std::string k, v;
                if (!id(k)) return false;
                while (i < n && (s[i] == ' ' || s[i] == '\t')) i++;
                if (i >= n || s[i] != '=') return false;
                i++;
                while (i < n && (s[i] == ' ' || s[i] == '\t')) i++;
                if (!id(v)) return false;
                st.attrs.push_back(std::make_pair(k, v));
            }
        }
        out.push_back(st);
    }
}
}  // namespace orc_graph

ORC_GT bool read_graphviz(const std::string &dot, ORC_G &g, dynamic_properties &, const std::string & /*node_id*/) {
    std::vector<orc_graph::DotStmt> st;
    if (!orc_graph::dot_parse(dot, st)) return false;
    std::map<std::string, std::map<std::string, std::string> > nodes;  // sorted by id, like read_graphviz_new's parser_result
    for (std::size_t i = 0; i < st.size(); i++) {
        nodes[st[i].a];
        if (!st[i].b.empty()) nodes[st[i].b];
        else for (std::size_t j = 0; j < st[i].attrs.size(); j++) nodes[st[i].a][st[i].attrs[j].first] = st[i].attrs[j].second;
    }
    std::map<std::string, std::size_t> index;
    for (std::map<std::string, std::map<std::string, std::string> >::iterator it = nodes.begin(); it != nodes.end(); ++it) {
        orc_graph::VertexRec r;
        r.name = it->first;
        if (it->second.count("label")) r.color = it->second["label"];
        index[it->first] = g.v.size();
        g.v.push_back(r);
    }
    for (std::size_t i = 0; i < st.size(); i++) {
        if (st[i].b.empty()) continue;
        int w = 0;
        for (std::size_t j = 0; j < st[i].attrs.size(); j++)
            if (st[i].attrs[j].first == "coverage") w = atoi(st[i].attrs[j].second.c_str());
        orc_graph::OutEdge e;
        e.target = index[st[i].b];
        e.id = g.edge_weight_store.size();
        g.edge_weight_store.push_back(w);
        g.v[index[st[i].a]].out.push_back(e);
    }
    return true;
}
#undef ORC_G
#undef ORC_GT

/* Boost.Regex subset on std::regex (POSIX extended) */
class regex {
  public:
    enum flag_type { extended = 1 };
    std::regex re;
    /* Boost's POSIX-extended flavour accepts the \\s / \\S escapes; std::regex::extended does not */
    static std::string posix(const char *p) {
        std::string o;
        for (; *p; p++) {
            if (p[0] == '\\' && p[1] == 's') { o += "[[:space:]]"; p++; }
            else if (p[0] == '\\' && p[1] == 'S') { o += "[^[:space:]]"; p++; }
            else o += *p;
        }
        return o;
    }
    regex(const char *p, int = 0) : re(posix(p), std::regex::extended) {}
};
template <class It> struct match_results : std::match_results<It> {};
inline bool regex_search(const char *s, match_results<const char *> &m, const regex &r) {
    /* "^...$" must match a LINE of the DOT text (Boost's default: ^ and $ match at embedded newlines) */
    const char *p = s;
    while (*p) {
        const char *e = p;
        while (*e && *e != '\n') e++;
        if (std::regex_search(p, e, static_cast<std::match_results<const char *> &>(m), r.re)) return true;
        p = *e ? e + 1 : e;
    }
    return false;
}

/* Boost.StringAlgo subset */
struct orc_any_of { std::string set; };
inline orc_any_of is_any_of(const char *s) { orc_any_of a; a.set = s; return a; }
template <class Vec> inline Vec &split(Vec &out, const std::string &in, const orc_any_of &pred) {
    out.clear();
    std::string cur;
    for (std::size_t i = 0; i < in.size(); i++) {
        if (pred.set.find(in[i]) != std::string::npos) { out.push_back(cur); cur.clear(); }
        else cur += in[i];
    }
    out.push_back(cur);
    return out;
}

}  // namespace boost
#endif
