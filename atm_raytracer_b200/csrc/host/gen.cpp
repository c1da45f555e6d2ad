// gen.cpp -- the thin C++ host of the `gen` subcommand (generator/mod.rs:47-99): YAML config + CLI
// overrides (generator/params.rs:447-777) -> flat PODs of include/atmrt.h -> the CUDA library ->
// PNG (+ optional per-pixel metadata sidecar). The reference's host is Rust; there is no Rust
// toolchain in this image, so the same surface is restated in C++ on top of the same C ABI a Rust
// host would bind. Nothing here computes the hot path.
#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "atmrt_host.h"
#include "pgzip.h"
#include "../../../include/atmrt_fmt.h"

namespace atmrt_host {
extern std::string g_error;
int fail(int code, const std::string& msg);

// ---------------------------------------------------------------------------------------------
// A YAML subset big enough for the reference's Config (README.md:76-324): block maps and lists,
// flow maps/lists, scalars, comments, quoted strings. serde's externally tagged enums appear as a
// bare string (unit variant) or a single-key map.
// ---------------------------------------------------------------------------------------------
struct Node {
    enum Kind { Null, Scalar, Map, Seq } kind = Null;
    std::string scalar;
    std::vector<std::pair<std::string, Node>> map;
    std::vector<Node> seq;

    const Node* get(const std::string& key) const {
        if (kind != Map) return nullptr;
        for (const auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_null() const { return kind == Null || (kind == Scalar && (scalar == "~" || scalar == "null" || scalar.empty())); }
    double as_double(const std::string& what) const {
        if (kind != Scalar) throw std::runtime_error(what + ": expected a number");
        char* end = nullptr;
        double v = strtod(scalar.c_str(), &end);
        if (end == scalar.c_str() || *end != '\0') throw std::runtime_error(what + ": invalid number '" + scalar + "'");
        return v;
    }
    bool as_bool(const std::string& what) const {
        if (kind == Scalar && (scalar == "true" || scalar == "True")) return true;
        if (kind == Scalar && (scalar == "false" || scalar == "False")) return false;
        throw std::runtime_error(what + ": expected true/false");
    }
};

struct Line {
    int indent;
    std::string text;  // without indentation and trailing comment
    int number;
};

static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

static std::string strip_comment(const std::string& s) {
    bool sq = false, dq = false;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        if (c == '"' && !sq) dq = !dq;
        if (c == '#' && !sq && !dq && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}

static std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
    return s;
}

// flow-style value: {a: 1, b: [2, 3]} / [..] / scalar
struct FlowParser {
    const std::string& s;
    size_t i = 0;
    explicit FlowParser(const std::string& str) : s(str) {}
    void ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i;
    }
    Node value() {
        ws();
        Node n;
        if (i < s.size() && s[i] == '{') {
            ++i;
            n.kind = Node::Map;
            ws();
            if (i < s.size() && s[i] == '}') {
                ++i;
                return n;
            }
            for (;;) {
                ws();
                std::string key = token(":");
                ws();
                if (i >= s.size() || s[i] != ':') throw std::runtime_error("flow map: expected ':' after '" + key + "'");
                ++i;
                Node v = value();
                n.map.emplace_back(unquote(trim(key)), v);
                ws();
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == '}') {
                    ++i;
                    return n;
                }
                throw std::runtime_error("flow map: expected ',' or '}'");
            }
        }
        if (i < s.size() && s[i] == '[') {
            ++i;
            n.kind = Node::Seq;
            ws();
            if (i < s.size() && s[i] == ']') {
                ++i;
                return n;
            }
            for (;;) {
                n.seq.push_back(value());
                ws();
                if (i < s.size() && s[i] == ',') {
                    ++i;
                    continue;
                }
                if (i < s.size() && s[i] == ']') {
                    ++i;
                    return n;
                }
                throw std::runtime_error("flow list: expected ',' or ']'");
            }
        }
        std::string t = trim(token(",}]"));
        if (t.empty() || t == "~" || t == "null") return n;
        n.kind = Node::Scalar;
        n.scalar = unquote(t);
        return n;
    }
    std::string token(const char* stops) {
        size_t a = i;
        bool sq = false, dq = false;
        while (i < s.size()) {
            char c = s[i];
            if (c == '\'' && !dq) sq = !sq;
            if (c == '"' && !sq) dq = !dq;
            if (!sq && !dq && strchr(stops, c)) break;
            ++i;
        }
        return s.substr(a, i - a);
    }
};

static Node parse_inline(const std::string& text) {
    std::string t = trim(text);
    FlowParser fp(t);
    Node n = fp.value();
    fp.ws();
    if (fp.i != t.size()) {  // a plain scalar containing ',' etc.
        Node s;
        s.kind = Node::Scalar;
        s.scalar = unquote(t);
        return s;
    }
    return n;
}

// position of the ':' that separates key and value in a block-map line (or npos)
static size_t key_colon(const std::string& t) {
    bool sq = false, dq = false;
    int depth = 0;
    for (size_t i = 0; i < t.size(); ++i) {
        char c = t[i];
        if (c == '\'' && !dq) sq = !sq;
        if (c == '"' && !sq) dq = !dq;
        if (sq || dq) continue;
        if (c == '{' || c == '[') ++depth;
        if (c == '}' || c == ']') --depth;
        if (c == ':' && depth == 0 && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
    }
    return std::string::npos;
}

struct BlockParser {
    std::vector<Line> lines;
    size_t pos = 0;

    Node block(int indent) {
        Node n;
        if (pos >= lines.size() || lines[pos].indent < indent) return n;
        const int ind = lines[pos].indent;
        if (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-") {
            n.kind = Node::Seq;
            while (pos < lines.size() && lines[pos].indent == ind && (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-")) {
                std::string rest = lines[pos].text.size() > 2 ? trim(lines[pos].text.substr(2)) : std::string();
                const int child = ind + 2;
                if (rest.empty()) {
                    ++pos;
                    n.seq.push_back(block(ind + 1));
                } else if (key_colon(rest) != std::string::npos && rest[0] != '{' && rest[0] != '[') {
                    // "- key: value" starts a map whose other keys follow at the same column as `key`
                    lines[pos].indent = child;
                    lines[pos].text = rest;
                    n.seq.push_back(block(child));
                } else if (rest.rfind("- ", 0) == 0) {
                    // "- - value" starts a nested list whose other items follow at the column of the inner dash
                    lines[pos].indent = child;
                    lines[pos].text = rest;
                    n.seq.push_back(block(child));
                } else {
                    ++pos;
                    n.seq.push_back(parse_inline(rest));
                }
            }
            return n;
        }
        n.kind = Node::Map;
        while (pos < lines.size() && lines[pos].indent == ind) {
            const std::string& t = lines[pos].text;
            size_t c = key_colon(t);
            if (c == std::string::npos) throw std::runtime_error("line " + std::to_string(lines[pos].number) + ": expected 'key: value'");
            std::string key = unquote(trim(t.substr(0, c)));
            std::string rest = trim(t.substr(c + 1));
            ++pos;
            if (rest.empty()) {
                if (pos < lines.size() && lines[pos].indent > ind)
                    n.map.emplace_back(key, block(lines[pos].indent));
                else if (pos < lines.size() && lines[pos].indent == ind && lines[pos].text.rfind("- ", 0) == 0)
                    n.map.emplace_back(key, block(ind));  // list at the same indentation as its key
                else
                    n.map.emplace_back(key, Node());
            } else {
                n.map.emplace_back(key, parse_inline(rest));
            }
        }
        if (pos < lines.size() && lines[pos].indent > ind) throw std::runtime_error("line " + std::to_string(lines[pos].number) + ": bad indentation");
        return n;
    }
};

static Node parse_yaml(const std::string& text) {
    BlockParser bp;
    size_t a = 0;
    int number = 0;
    while (a <= text.size()) {
        size_t b = text.find('\n', a);
        if (b == std::string::npos) b = text.size();
        std::string raw = text.substr(a, b - a);
        a = b + 1;
        ++number;
        std::string nc = strip_comment(raw);
        std::string t = trim(nc);
        if (t.empty() || t == "---") continue;
        int indent = 0;
        while (indent < (int)nc.size() && nc[indent] == ' ') ++indent;
        bp.lines.push_back({indent, t, number});
    }
    if (bp.lines.empty()) return Node();
    Node n = bp.block(bp.lines[0].indent);
    if (bp.pos != bp.lines.size()) throw std::runtime_error("line " + std::to_string(bp.lines[bp.pos].number) + ": unexpected content");
    return n;
}

// ---------------------------------------------------------------------------------------------
// Config (generator/params.rs:447-494) with the reference's defaults
// ---------------------------------------------------------------------------------------------
struct ConfObject {
    atmrt_object o{};
    std::string texture_path;
};

struct Config {
    std::string terrain_folder = "./terrain";
    std::vector<ConfObject> objects;
    double terrain_alpha = 1.0;
    double latitude = 0.0, longitude = 0.0;
    atmrt_altitude altitude{ATMRT_ALT_RELATIVE, 0, 1.0};
    double direction = 0.0, tilt = 0.0, fov = 30.0, max_distance = 150000.0;
    int coloring = ATMRT_COLORING_SHADING, palette = ATMRT_PALETTE_IMPROVED;
    double water_level = 0.0, ambient_light = 0.4, light_zenith_angle = 45.0, light_dir = 0.0;
    bool fog = false;
    double fog_distance = 0.0;
    atmrt_atmosphere_def atmosphere{};
    int earth_model = ATMRT_EARTH_SPHERICAL;
    double radius = 6371000.0;
    double ellipsoid_b = 0.0;
    int generator = ATMRT_GENERATOR_FAST;
    double wavelength = 530e-9;
    bool straight_rays = false;
    double simulation_step = 50.0;
    std::string file = "./output.png", file_metadata;
    int width = 640, height = 480;
    int gpus = 1;  // --gpus N (not a flag of the reference): column blocks over N GPUs of this box
    // the overlays of renderer::output_image (params.rs:403-410)
    std::vector<atmrt_host_tick> ticks, vertical_ticks;
    bool show_eye_level = false, show_flat_horizon = false;
};

static atmrt_atmosphere_def us_76() {  // AtmosphereDef::us_76() (params.rs:453)
    atmrt_atmosphere_def a{};
    a.pressure_altitude = 0.0, a.pressure = 101325.0, a.temperature_altitude = 0.0, a.temperature = 288.15, a.humidity = 0.0;
    const double starts[] = {0.0, 11000.0, 20000.0, 32000.0, 47000.0, 51000.0, 71000.0};
    const double grads[] = {-0.0065, 0.0, 0.001, 0.0028, 0.0, -0.0028, -0.002};
    a.n_functions = 7;
    for (int i = 0; i < 7; ++i) a.fn_start_altitude[i] = starts[i], a.fn_gradient[i] = grads[i];
    return a;
}

static void tagged(const Node& n, const std::string& what, std::string* tag, const Node** body) {
    static const Node empty;
    if (n.kind == Node::Scalar) {
        *tag = n.scalar;
        *body = &empty;
        return;
    }
    if (n.kind == Node::Map && n.map.size() == 1) {
        *tag = n.map[0].first;
        *body = &n.map[0].second;
        return;
    }
    throw std::runtime_error("invalid " + what);
}

static atmrt_altitude parse_altitude(const Node& n) {
    std::string tag;
    const Node* body;
    tagged(n, "altitude", &tag, &body);
    atmrt_altitude a{};
    if (tag == "Absolute")
        a.kind = ATMRT_ALT_ABSOLUTE;
    else if (tag == "Relative")
        a.kind = ATMRT_ALT_RELATIVE;
    else
        throw std::runtime_error("unknown altitude kind " + tag);
    a.value = body->as_double("altitude");
    return a;
}

static double num(const Node& m, const char* key, double def) {
    const Node* n = m.get(key);
    return n && !n->is_null() ? n->as_double(key) : def;
}

static void parse_position(const Node& n, double* lat, double* lon, atmrt_altitude* alt) {
    *lat = num(n, "latitude", 0.0);
    *lon = num(n, "longitude", 0.0);
    if (const Node* a = n.get("altitude")) *alt = parse_altitude(*a);
}

// Vec<Tick> / Vec<VerticalTick> (params.rs:325-385): externally tagged Single { azimuth | elevation, size, labelled } or
// Multiple { bias, step, size, labelled }; serde has no defaults for these fields
static std::vector<atmrt_host_tick> parse_ticks(const Node& list, const std::string& what, const char* single_key) {
    std::vector<atmrt_host_tick> out;
    if (list.is_null()) return out;
    if (list.kind != Node::Seq) throw std::runtime_error("output." + what + ": expected a list");
    for (const Node& n : list.seq) {
        std::string tag;
        const Node* body;
        tagged(n, "output." + what + " entry", &tag, &body);
        auto field = [&](const char* key) -> const Node& {
            const Node* f = body->get(key);
            if (!f || f->is_null()) throw std::runtime_error("output." + what + ": " + tag + " is missing field `" + key + "`");
            return *f;
        };
        atmrt_host_tick t{};
        if (tag == "Single")
            t.multiple = 0, t.angle = field(single_key).as_double(single_key);
        else if (tag == "Multiple")
            t.multiple = 1, t.bias = field("bias").as_double("bias"), t.step = field("step").as_double("step");
        else
            throw std::runtime_error("output." + what + ": unknown variant " + tag + " (Single, Multiple)");
        const double size = field("size").as_double("size");
        if (!(size >= 0.0) || size > 4294967295.0 || size != std::floor(size)) throw std::runtime_error("output." + what + ": size must be a u32");
        t.size = (uint32_t)size;
        t.labelled = field("labelled").as_bool("labelled") ? 1 : 0;
        out.push_back(t);
    }
    return out;
}

static void apply_yaml(const Node& doc, Config* c) {
    if (doc.kind == Node::Null) return;
    if (doc.kind != Node::Map) throw std::runtime_error("config: top level must be a map");
    if (const Node* scene = doc.get("scene")) {
        if (const Node* t = scene->get("terrain_folder")) c->terrain_folder = t->scalar;
        c->terrain_alpha = num(*scene, "terrain_alpha", c->terrain_alpha);
        if (const Node* objs = scene->get("objects")) {
            if (objs->kind != Node::Seq && !objs->is_null()) throw std::runtime_error("scene.objects must be a list");
            for (const Node& on : objs->seq) {
                ConfObject co;
                co.o.altitude = atmrt_altitude{ATMRT_ALT_RELATIVE, 0, 1.0};
                const Node* pos = on.get("position");
                const Node* shape = on.get("shape");
                if (!pos || !shape) throw std::runtime_error("object needs position and shape");
                parse_position(*pos, &co.o.latitude, &co.o.longitude, &co.o.altitude);
                co.o.color[0] = co.o.color[1] = co.o.color[2] = co.o.color[3] = 1.0;
                if (!on.get("color")) throw std::runtime_error("object needs a color");  // `color` has no serde default (object/mod.rs:158-163)
                if (const Node* col = on.get("color")) {
                    co.o.color[0] = num(*col, "r", 1.0), co.o.color[1] = num(*col, "g", 1.0), co.o.color[2] = num(*col, "b", 1.0);
                    co.o.color[3] = num(*col, "a", 1.0);  // default_alpha, object/mod.rs:147-156
                }
                std::string tag;
                const Node* body;
                tagged(*shape, "shape", &tag, &body);
                if (tag == "Cylinder") {  // ConfShape::into_shape, object/mod.rs:41-54
                    co.o.kind = ATMRT_OBJECT_FRUSTUM;
                    co.o.r1 = co.o.r2 = num(*body, "radius", 0.0);
                    co.o.height = num(*body, "height", 0.0);
                } else if (tag == "Cone") {
                    co.o.kind = ATMRT_OBJECT_FRUSTUM;
                    co.o.r1 = num(*body, "radius", 0.0), co.o.r2 = 0.0, co.o.height = num(*body, "height", 0.0);
                } else if (tag == "Frustum") {
                    co.o.kind = ATMRT_OBJECT_FRUSTUM;
                    co.o.r1 = num(*body, "r1", 0.0), co.o.r2 = num(*body, "r2", 0.0), co.o.height = num(*body, "height", 0.0);
                } else if (tag == "Billboard") {
                    co.o.kind = ATMRT_OBJECT_BILLBOARD;
                    co.o.width = num(*body, "width", 0.0), co.o.height = num(*body, "height", 0.0);
                    const Node* tp = body->get("texture_path");
                    if (!tp) throw std::runtime_error("Billboard needs texture_path");
                    co.texture_path = tp->scalar;
                } else {
                    throw std::runtime_error("unknown shape " + tag);
                }
                c->objects.push_back(co);
            }
        }
    }
    if (const Node* view = doc.get("view")) {
        if (const Node* pos = view->get("position")) parse_position(*pos, &c->latitude, &c->longitude, &c->altitude);
        if (const Node* fr = view->get("frame")) {
            c->direction = num(*fr, "direction", c->direction), c->tilt = num(*fr, "tilt", c->tilt);
            c->fov = num(*fr, "fov", c->fov), c->max_distance = num(*fr, "max_distance", c->max_distance);
        }
        if (const Node* col = view->get("coloring")) {
            std::string tag;
            const Node* body;
            tagged(*col, "coloring", &tag, &body);
            c->water_level = num(*body, "water_level", 0.0);
            if (tag == "Simple") {
                c->coloring = ATMRT_COLORING_SIMPLE;
            } else if (tag == "Shading") {
                c->coloring = ATMRT_COLORING_SHADING;
                c->ambient_light = num(*body, "ambient_light", 0.4);
                c->light_zenith_angle = num(*body, "light_zenith_angle", 45.0);
                c->light_dir = num(*body, "light_dir", 0.0);
                if (const Node* pal = body->get("palette")) {
                    if (pal->scalar == "Legacy")
                        c->palette = ATMRT_PALETTE_LEGACY;
                    else if (pal->scalar == "Improved")
                        c->palette = ATMRT_PALETTE_IMPROVED;
                    else
                        throw std::runtime_error("unknown palette " + pal->scalar);
                }
            } else {
                throw std::runtime_error("unknown coloring " + tag);
            }
        }
        if (const Node* fog = view->get("fog_distance")) {
            c->fog = !fog->is_null();
            if (c->fog) c->fog_distance = fog->as_double("fog_distance");
        }
    }
    if (const Node* atm = doc.get("atmosphere")) {
        if (!atm->is_null()) {
            atmrt_atmosphere_def a{};
            const Node* pr = atm->get("pressure");
            const Node* first = atm->get("first_temperature_function");
            if (!pr || !first) throw std::runtime_error("atmosphere needs pressure and first_temperature_function");
            a.pressure_altitude = num(*pr, "altitude", 0.0), a.pressure = num(*pr, "pressure", 101325.0);
            a.humidity = num(*atm, "humidity", 0.0);
            std::vector<std::pair<double, const Node*>> fns{{0.0, first}};
            if (const Node* next = atm->get("next_functions"))
                for (const Node& nf : next->seq) {
                    const Node* fn = nf.get("function");
                    if (!fn) throw std::runtime_error("next_functions entry needs altitude and function");
                    fns.emplace_back(num(nf, "altitude", 0.0), fn);
                }
            if ((int)fns.size() > ATMRT_MAX_ATM_FUNCTIONS) throw std::runtime_error("too many temperature functions");
            a.n_functions = (int)fns.size();
            for (size_t i = 0; i < fns.size(); ++i) {
                std::string tag;
                const Node* body;
                tagged(*fns[i].second, "temperature function", &tag, &body);
                a.fn_start_altitude[i] = fns[i].first;
                if (tag == "Linear") {
                    a.fn_kind[i] = ATMRT_FUNCTION_LINEAR;
                    a.fn_gradient[i] = num(*body, "gradient", 0.0);
                    continue;
                }
                if (tag != "Spline") throw std::runtime_error("unknown temperature function '" + tag + "'");
                a.fn_kind[i] = ATMRT_FUNCTION_SPLINE;
                a.fn_boundary[i] = ATMRT_SPLINE_NATURAL;
                if (const Node* bc = body->get("boundary_condition")) {
                    std::string bc_tag;
                    const Node* bc_body;
                    tagged(*bc, "boundary_condition", &bc_tag, &bc_body);
                    if (bc_tag == "Derivatives")
                        a.fn_boundary[i] = ATMRT_SPLINE_DERIVATIVES;
                    else if (bc_tag == "SecondDerivatives")
                        a.fn_boundary[i] = ATMRT_SPLINE_SECOND_DERIVATIVES;
                    else if (bc_tag != "Natural")
                        throw std::runtime_error("unknown Spline boundary condition '" + bc_tag + "'");
                    if (bc_tag != "Natural") {
                        if (bc_body->kind != Node::Seq || bc_body->seq.size() != 2) throw std::runtime_error(bc_tag + " needs two numbers");
                        a.fn_boundary_values[i][0] = bc_body->seq[0].as_double(bc_tag), a.fn_boundary_values[i][1] = bc_body->seq[1].as_double(bc_tag);
                    }
                }
                const Node* pts = body->get("points");
                if (!pts || pts->kind != Node::Seq || pts->seq.size() < 2) throw std::runtime_error("a Spline needs at least two points");
                if (a.n_spline_points + (int)pts->seq.size() > ATMRT_MAX_SPLINE_POINTS) throw std::runtime_error("too many Spline points");
                a.fn_first_point[i] = a.n_spline_points, a.fn_n_points[i] = (int)pts->seq.size();
                for (const Node& pt : pts->seq) {
                    if (pt.kind != Node::Seq || pt.seq.size() != 2) throw std::runtime_error("a Spline point is a pair (altitude, temperature)");
                    a.spline_points[a.n_spline_points][0] = pt.seq[0].as_double("Spline point altitude");
                    a.spline_points[a.n_spline_points][1] = pt.seq[1].as_double("Spline point temperature");
                    ++a.n_spline_points;
                }
            }
            const Node* tfp = atm->get("temperature_fixed_point");
            if (!tfp && a.n_spline_points == 0) throw std::runtime_error("temperature_fixed_point is required when every function is Linear");
            if (tfp) a.temperature_altitude = num(*tfp, "altitude", 0.0), a.temperature = num(*tfp, "temperature", 288.15);  // ignored beside a Spline (README.md:318-323)
            c->atmosphere = a;
        }
    }
    if (const Node* es = doc.get("earth_shape")) {
        std::string tag;
        const Node* body;
        tagged(*es, "earth_shape", &tag, &body);
        if (tag == "Spherical")
            c->earth_model = ATMRT_EARTH_SPHERICAL, c->radius = num(*body, "radius", 6371000.0);
        else if (tag == "SimpleSphere")
            c->earth_model = ATMRT_EARTH_SPHERICAL, c->radius = 6371000.0;
        else if (tag == "FlatDistorted")
            c->earth_model = ATMRT_EARTH_FLAT_DISTORTED;
        else if (tag == "Ellipsoid")
            c->earth_model = ATMRT_EARTH_ELLIPSOID, c->radius = num(*body, "a", 6378137.0), c->ellipsoid_b = num(*body, "b", 6356752.314245);
        else if (tag == "Wgs84")  // earth_model/mod.rs:15-16, 66-70
            c->earth_model = ATMRT_EARTH_ELLIPSOID, c->radius = 6378137.0, c->ellipsoid_b = 6356752.314245;
        else if (tag == "AzimuthalEquidistant")
            c->earth_model = ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT;
        else if (tag == "ObserverAe")
            c->earth_model = ATMRT_EARTH_OBSERVER_AE, c->radius = num(*body, "proj_radius", 6371000.0);
        else if (tag == "SimpleObserverAe")  // mod.rs:139-143
            c->earth_model = ATMRT_EARTH_OBSERVER_AE, c->radius = 6371000.0;
        else
            throw std::runtime_error("unknown earth_shape " + tag);
    }
    c->wavelength = num(doc, "wavelength", c->wavelength);
    if (const Node* s = doc.get("straight_rays")) c->straight_rays = s->as_bool("straight_rays");
    c->simulation_step = num(doc, "simulation_step", c->simulation_step);
    if (const Node* out = doc.get("output")) {
        if (const Node* f = out->get("file")) c->file = f->scalar;
        if (const Node* f = out->get("file_metadata"))
            if (!f->is_null()) c->file_metadata = f->scalar;
        c->width = (int)num(*out, "width", c->width), c->height = (int)num(*out, "height", c->height);
        if (const Node* g = out->get("generator"))
        {
            if (g->scalar == "Fast") c->generator = ATMRT_GENERATOR_FAST;
            else if (g->scalar == "Rectilinear") c->generator = ATMRT_GENERATOR_RECTILINEAR;
            else if (g->scalar == "InterpolatingRectilinear") c->generator = ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR;
            else throw std::runtime_error("unknown generator " + g->scalar + " (Fast, Rectilinear, InterpolatingRectilinear)");
        }
        if (const Node* t = out->get("ticks")) c->ticks = parse_ticks(*t, "ticks", "azimuth");
        if (const Node* t = out->get("vertical_ticks")) c->vertical_ticks = parse_ticks(*t, "vertical_ticks", "elevation");
        if (const Node* t = out->get("show_eye_level")) c->show_eye_level = t->as_bool("show_eye_level");
        if (const Node* t = out->get("show_flat_horizon")) c->show_flat_horizon = t->as_bool("show_flat_horizon");
    }
}

static std::string read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::string s;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

// read_config (params.rs:694-777): YAML first, then the CLI flags.
static Config read_config(int argc, const char* const* argv) {
    Config c;
    c.atmosphere = us_76();
    std::map<std::string, std::string> val;
    std::map<std::string, bool> flag;
    const std::map<std::string, std::string> takes = {
        {"-c", "config"}, {"--config", "config"}, {"-t", "terrain"}, {"--terrain", "terrain"}, {"-l", "lat"}, {"--lat", "lat"},
        {"-g", "lon"}, {"--lon", "lon"}, {"-a", "alt"}, {"--alt", "alt"}, {"-e", "elev"}, {"--elev", "elev"}, {"-d", "dir"},
        {"--dir", "dir"}, {"-f", "fov"}, {"--fov", "fov"}, {"-i", "tilt"}, {"--tilt", "tilt"}, {"-m", "maxdist"},
        {"--maxdist", "maxdist"}, {"--step", "step"}, {"-R", "radius"}, {"--radius", "radius"}, {"--output", "output"},
        {"--output-meta", "output-meta"}, {"-w", "width"}, {"--width", "width"}, {"-h", "height"}, {"--height", "height"}, {"--gpus", "gpus"}};
    const std::map<std::string, std::string> flags = {{"--flat", "flat"}, {"-s", "straight"}, {"--straight", "straight"}};
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        auto t = takes.find(a);
        if (t != takes.end()) {
            if (i + 1 >= argc) throw std::runtime_error("missing value for " + a);
            val[t->second] = argv[++i];  // AllowLeadingHyphen: negative numbers are values (params.rs:533)
            continue;
        }
        auto f = flags.find(a);
        if (f != flags.end()) {
            flag[f->second] = true;
            continue;
        }
        throw std::runtime_error("unknown argument " + a);
    }
    if (val.count("alt") && val.count("elev")) throw std::runtime_error("--alt conflicts with --elev");
    if (flag.count("flat") && val.count("radius")) throw std::runtime_error("--flat conflicts with --radius");
    if (val.count("config")) apply_yaml(parse_yaml(read_file(val["config"])), &c);
    auto d = [&](const char* k) { return strtod(val[k].c_str(), nullptr); };
    if (val.count("terrain")) c.terrain_folder = val["terrain"];
    if (val.count("output")) c.file = val["output"];
    if (val.count("output-meta")) c.file_metadata = val["output-meta"];
    if (val.count("width")) c.width = atoi(val["width"].c_str());
    if (val.count("height")) c.height = atoi(val["height"].c_str());
    if (val.count("lat")) c.latitude = d("lat");
    if (val.count("lon")) c.longitude = d("lon");
    if (val.count("alt")) c.altitude = atmrt_altitude{ATMRT_ALT_ABSOLUTE, 0, d("alt")};
    if (val.count("elev")) c.altitude = atmrt_altitude{ATMRT_ALT_RELATIVE, 0, d("elev")};
    if (val.count("dir")) c.direction = d("dir");
    if (val.count("fov")) c.fov = d("fov");
    if (val.count("tilt")) c.tilt = d("tilt");
    if (val.count("maxdist")) c.max_distance = d("maxdist") * 1e3;  // km on the CLI (params.rs:751-754)
    if (val.count("step")) c.simulation_step = d("step");
    if (flag.count("flat")) c.earth_model = ATMRT_EARTH_FLAT_DISTORTED;
    if (val.count("radius")) c.earth_model = ATMRT_EARTH_SPHERICAL, c.radius = d("radius") * 1e3;  // km (params.rs:764-767)
    if (flag.count("straight")) c.straight_rays = true;
    if (val.count("gpus")) {
        c.gpus = atoi(val["gpus"].c_str());
        if (c.gpus < 1) throw std::runtime_error("--gpus must be at least 1");
    }
    return c;
}

// ConfColoring::into_coloring light vector (params.rs:243-259); host libm like the reference.
static void light_direction(const Config& c, double out[3]) {
    const double PI = 3.14159265358979323846;
    auto rad = [&](double d) { return d * (PI / 180.0); };
    double zen = rad(c.light_zenith_angle), ld = rad(c.light_dir);
    double lon = rad(c.longitude), lat = rad(c.latitude);
    double sinlon = std::sin(lon), coslon = std::cos(lon);
    double north[3], east[3], up[3];
    if (c.earth_model == ATMRT_EARTH_FLAT_DISTORTED || c.earth_model == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT || c.earth_model == ATMRT_EARTH_OBSERVER_AE) {
        north[0] = -coslon, north[1] = -sinlon, north[2] = 0.0;
        east[0] = -sinlon, east[1] = coslon, east[2] = 0.0;
        up[0] = 0.0, up[1] = 0.0, up[2] = 1.0;
    } else {
        double sinlat = std::sin(lat), coslat = std::cos(lat);
        up[0] = coslat * coslon, up[1] = coslat * sinlon, up[2] = sinlat;
        north[0] = -sinlat * coslon, north[1] = -sinlat * sinlon, north[2] = coslat;
        east[0] = -sinlon, east[1] = coslon, east[2] = 0.0;
    }
    double az = rad(c.direction);
    double v[3];
    for (int i = 0; i < 3; ++i) {
        double front = north[i] * std::cos(az) + east[i] * std::sin(az);
        double right = east[i] * std::cos(az) - north[i] * std::sin(az);
        v[i] = -front * std::sin(zen) * std::cos(ld) + right * std::sin(zen) * std::sin(ld) + up[i] * std::cos(zen);
    }
    double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int i = 0; i < 3; ++i) out[i] = v[i] / n;
}

static atmrt_params into_params(const Config& c) {
    atmrt_params p{};
    p.latitude = c.latitude, p.longitude = c.longitude, p.altitude = c.altitude;
    p.direction = c.direction, p.tilt = c.tilt, p.fov = c.fov, p.max_distance = c.max_distance;
    p.earth_model = c.earth_model, p.straight_rays = c.straight_rays ? 1 : 0, p.radius = c.radius, p.ellipsoid_b = c.ellipsoid_b;
    p.wavelength = c.wavelength, p.simulation_step = c.simulation_step, p.atmosphere = c.atmosphere;
    p.terrain_alpha = c.terrain_alpha;
    p.coloring = c.coloring, p.water_level = c.water_level;
    if (c.coloring == ATMRT_COLORING_SHADING) {  // Coloring::Simple carries none of these (params.rs:215-227)
        p.palette = c.palette, p.ambient_light = c.ambient_light;
        light_direction(c, p.light_dir);
    }
    if (c.earth_model == ATMRT_EARTH_FLAT_DISTORTED || c.earth_model == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT) p.radius = 0.0;
    if (c.earth_model != ATMRT_EARTH_ELLIPSOID) p.ellipsoid_b = 0.0;
    p.simple_max_distance = c.max_distance;
    p.fog_enabled = c.fog ? 1 : 0, p.fog_distance = c.fog_distance;
    p.generator = c.generator;
    p.width = c.width, p.height = c.height, p.x0 = 0, p.x1 = c.width;
    return p;
}

static bool write_metadata(const std::string& path, const atmrt_params& p, const atmrt_meta* meta, size_t npix, const std::vector<double>& elevation_angle,
                           const std::vector<double>& azimuth, const std::vector<int32_t>* counts, const std::vector<atmrt_trace_point>* points,
                           int max_points) {
    // NOT the reference's file: generator/mod.rs:26-45 writes gzip(bincode(AllData{params, result})), and `params` holds
    // atm_refraction's lowered Atmosphere and Environment, whose serde layout lives in an un-vendored crate (SURVEY section
    // 8 f2) -- `view` cannot read what is written here. Own documented sidecar: gzip of
    //   "ATMRTMETA2\n" (or "ATMRTMETA3\n" with the trace lists below), i32 width, i32 height, i32 generator (0 Fast,
    //   1 Rectilinear, 2 InterpolatingRectilinear), i32 0,
    //   ResultPixel.elevation_angle: Fast f64[height] (one per row), else f64[height][width]
    //   ResultPixel.azimuth:         Fast f64[width] (one per column, wrapped into [0, 360)), else f64[height][width]
    //   width * height records of 4 little-endian f64: lat, lon, elevation, distance of the FIRST trace point (NaN = none)
    // version 3 (scenes in which a pixel can hold more than one trace point: translucent terrain, objects), after that:
    //   i32 max_points, i32 counts[height][width] (true counts), then per pixel min(count, max_points) records of
    //   atmrt_trace_point (include/atmrt.h: lat, lon, distance, elevation, path_length, normal[3], color[4] as f64,
    //   i32 is_terrain, i32 step) -- ResultPixel.trace_points (generators/mod.rs:18-30)
    // (a series of gzip members compressed on all host threads -- pgzip.h; any gzip reader sees one stream)
    ParallelGzip z(path);
    if (!z.ok()) return false;
    const bool lists = counts && points;
    const char* magic = lists ? "ATMRTMETA3\n" : "ATMRTMETA2\n";
    int32_t hdr[4] = {p.width, p.height, p.generator, 0};
    z.put(magic, 11), z.put(hdr, sizeof hdr);
    z.put(elevation_angle.data(), elevation_angle.size() * sizeof(double));
    z.put(azimuth.data(), azimuth.size() * sizeof(double));
    z.put(meta, npix * sizeof(atmrt_meta));
    if (lists) {
        const int32_t mp = max_points;
        z.put(&mp, sizeof mp);
        z.put(counts->data(), npix * sizeof(int32_t));
        for (size_t i = 0; i < npix && z.ok(); ++i) {
            const size_t n = (size_t)std::min<int32_t>((*counts)[i], max_points);
            if (n) z.put(points->data() + i * (size_t)max_points, n * sizeof(atmrt_trace_point));
        }
    }
    return z.close();
}

// Terrain::from_folder (terrain/mod.rs:66-83): every entry of the folder must be a terrain file; decoded on the host.
struct HostTerrain {
    std::vector<atmrt_tile_desc> descs;
    std::vector<int16_t*> posts;  // page-locked (atmrt_host_alloc): every GPU uploads its slice at the full rate of its link
    HostTerrain() = default;
    HostTerrain(const HostTerrain&) = delete;
    HostTerrain& operator=(const HostTerrain&) = delete;
    HostTerrain(HostTerrain&& o) noexcept : descs(std::move(o.descs)), posts(std::move(o.posts)) { o.posts.clear(); }
    ~HostTerrain() {
        for (int16_t* p : posts) atmrt_host_free(p);
    }
    std::vector<const int16_t*> ptrs() const { return std::vector<const int16_t*>(posts.begin(), posts.end()); }
};

static HostTerrain load_terrain_folder(const std::string& folder) {
    HostTerrain t;
    DIR* dir = opendir(folder.c_str());
    if (!dir) throw std::runtime_error("Error opening the terrain data directory " + folder);
    std::vector<std::string> names;
    while (dirent* e = readdir(dir)) {
        std::string n = e->d_name;
        if (n != "." && n != "..") names.push_back(n);
    }
    closedir(dir);
    std::sort(names.begin(), names.end());
    for (const std::string& n : names) {
        std::string path = folder + "/" + n;
        atmrt_tile_desc d{};
        // DTED by its header, else GeoTIFF by its name (Terrain::buffer_file, terrain/mod.rs:113-118)
        if (atmrt_host_read_tile(path.c_str(), &d, nullptr, 0) != 0) throw std::runtime_error("Could not buffer terrain file " + path + " (" + g_error + ")");
        const size_t np = (size_t)d.nlon * d.nlat;
        int16_t* buf = (int16_t*)atmrt_host_alloc(np * sizeof(int16_t));
        if (!buf) throw std::runtime_error("cannot allocate page-locked memory for " + path);
        t.descs.push_back(d);
        t.posts.push_back(buf);
        if (atmrt_host_read_tile(path.c_str(), &d, buf, np) != 0) throw std::runtime_error(g_error);
    }
    printf("Detected %zu terrain files\n", names.size());
    return t;
}

}  // namespace atmrt_host

using namespace atmrt_host;

extern "C" int atmrt_host_gen(int argc, const char* const* argv) {
    const auto start = std::chrono::steady_clock::now();
    auto t = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count(); };
    atmrt_group* group = nullptr;
    uint8_t* rgb = nullptr;
    atmrt_meta* meta = nullptr;
    try {
        Config c = read_config(argc, argv);
        printf("%.3f: Using terrain data directory: \"%s\"\n", t(), c.terrain_folder.c_str());
        HostTerrain terrain = load_terrain_folder(c.terrain_folder);

        atmrt_params p = into_params(c);
        std::vector<atmrt_object> objects;
        std::vector<std::vector<uint8_t>> textures(c.objects.size());
        std::vector<const uint8_t*> tex_ptrs(c.objects.size(), nullptr);
        for (size_t i = 0; i < c.objects.size(); ++i) {
            atmrt_object o = c.objects[i].o;
            if (o.kind == ATMRT_OBJECT_BILLBOARD) {  // texture path joined onto cwd (object/mod.rs:60-61)
                int w = 0, h = 0;
                if (atmrt_host_read_png(c.objects[i].texture_path.c_str(), nullptr, 0, &w, &h) != 0) throw std::runtime_error(g_error);
                textures[i].resize((size_t)w * h * 4);
                if (atmrt_host_read_png(c.objects[i].texture_path.c_str(), textures[i].data(), textures[i].size(), &w, &h) != 0)
                    throw std::runtime_error(g_error);
                o.texture_width = w, o.texture_height = h;
                tex_ptrs[i] = textures[i].data();
            }
            objects.push_back(o);
        }

        // One context per GPU, the panorama in column blocks (a group of one is the single-GPU render).
        auto check = [&](int rc, const char* what) {
            if (rc != 0) throw std::runtime_error(std::string(what) + ": " + atmrt_group_last_error(group));
        };
        check(atmrt_group_create(nullptr, c.gpus, &group), "atmrt_group_create");
        check(atmrt_group_set_params(group, &p), "atmrt_group_set_params");
        check(atmrt_group_set_objects(group, objects.data(), (int)objects.size(), tex_ptrs.data()), "atmrt_group_set_objects");
        printf("%.3f: Generating terrain cache...\n%.3f: Generating path cache...\n%.3f: Calculating pixels...\n", t(), t(), t());
        const size_t npix = (size_t)p.width * p.height;
        rgb = (uint8_t*)atmrt_host_alloc(npix * 3);
        if (!c.file_metadata.empty()) meta = (atmrt_meta*)atmrt_host_alloc(npix * sizeof(atmrt_meta));
        if (!rgb || (!c.file_metadata.empty() && !meta)) throw std::runtime_error("cannot allocate page-locked memory for the image");
        atmrt_stats st{};
        // the tiles go up, the image comes back: one call (the ray paths are integrated while the tiles are on their way)
        check(atmrt_group_render_tiles(group, terrain.descs.data(), (int)terrain.descs.size(), terrain.ptrs().data(), rgb, meta, nullptr, &st),
              "atmrt_group_render_tiles");
        printf("%.3f: Done calculating (terrain %.2f ms, paths %.2f ms, march %.2f ms on the GPU; %llu ray steps, %llu pixels hit)\n", t(),
               st.ms_terrain, st.ms_paths, st.ms_march, (unsigned long long)st.ray_steps, (unsigned long long)st.pixels_hit);
        printf("%.3f: Outputting image...\n", t());
        // ResultPixel.elevation_angle / azimuth (fast.rs:67-76, rectilinear.rs:78-116): the overlays and the sidecar read them
        const bool overlays = !c.ticks.empty() || !c.vertical_ticks.empty() || c.show_eye_level || c.show_flat_horizon;
        std::vector<double> el, az;
        if (overlays || !c.file_metadata.empty()) {
            el.resize(npix), az.resize(npix);
            check(atmrt_group_pixel_angles(group, el.data(), az.data()), "atmrt_group_pixel_angles");
        }
        if (overlays) {  // renderer::output_image (renderer/mod.rs:416-431): ticks, flat horizon, eye level over the picture
            atmrt_host_overlays ov{};
            ov.ticks = c.ticks.data(), ov.nticks = (int32_t)c.ticks.size();
            ov.vertical_ticks = c.vertical_ticks.data(), ov.nvertical_ticks = (int32_t)c.vertical_ticks.size();
            ov.direction = c.direction, ov.fov = c.fov, ov.tilt = c.tilt;
            ov.show_eye_level = c.show_eye_level;
            const bool flat_shape = c.earth_model == ATMRT_EARTH_FLAT_DISTORTED || c.earth_model == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT ||
                                    c.earth_model == ATMRT_EARTH_OBSERVER_AE;  // EarthModel::to_shape (earth_model/mod.rs:95-112)
            if (c.show_flat_horizon && flat_shape && !c.straight_rays) {
                atmrt_ctx* ctx0 = atmrt_group_context(group, 0);
                double alt = 0.0, n = 1.0;
                if (!ctx0 || atmrt_observer_altitude(ctx0, &alt) != 0 || atmrt_atmosphere_probe(ctx0, &alt, 1, nullptr, nullptr, &n) != 0)
                    throw std::runtime_error(std::string("show_flat_horizon: ") + (ctx0 ? atmrt_last_error(ctx0) : "no context"));
                ov.show_flat_horizon = 1;
                atmrt_host_flat_horizon_elevation(n, &ov.flat_horizon_elevation);
            }
            if (atmrt_host_draw_overlays(rgb, p.width, p.height, el.data(), az.data(), &ov) != 0) throw std::runtime_error(g_error);
            auto labelled = [](const atmrt_host_tick& tk) { return tk.labelled != 0; };
            if (std::any_of(c.ticks.begin(), c.ticks.end(), labelled) || std::any_of(c.vertical_ticks.begin(), c.vertical_ticks.end(), labelled))
                fprintf(stderr, "note: tick labels are drawn with a built-in bitmap face, not the reference's DejaVu Sans: label pixels differ\n");
        }
        if (atmrt_host_write_png(c.file.c_str(), rgb, p.width, p.height, 3) != 0) throw std::runtime_error(g_error);
        if (!c.file_metadata.empty()) {
            printf("%.3f: Outputting metadata...\n", t());
            if (p.generator == ATMRT_GENERATOR_FAST) {  // separable: one elevation per row, one azimuth per column
                std::vector<double> el_rows((size_t)p.height), az_cols(az.begin(), az.begin() + p.width);
                for (int y = 0; y < p.height; ++y) el_rows[(size_t)y] = el[(size_t)y * p.width];
                el.swap(el_rows), az.swap(az_cols);
            }
            // every trace point of every pixel where a pixel can hold more than its first (ResultPixel.trace_points)
            const bool lists = p.terrain_alpha != 1.0 || !objects.empty();
            const int max_points = 16;
            std::vector<int32_t> counts;
            std::vector<atmrt_trace_point> points;
            if (lists) {
                counts.resize(npix), points.resize(npix * (size_t)max_points);
                check(atmrt_group_render_trace(group, points.data(), counts.data(), max_points), "atmrt_group_render_trace");
                size_t clipped = 0;
                for (int32_t n : counts) clipped += n > max_points;
                if (clipped) fprintf(stderr, "warning: %zu pixels hold more than %d trace points; the sidecar keeps the first %d\n", clipped, max_points, max_points);
            }
            fprintf(stderr, "note: %s is this host's own sidecar layout (%s), not the reference's bincode container: `view` cannot read it\n",
                    c.file_metadata.c_str(), lists ? "ATMRTMETA3" : "ATMRTMETA2");
            if (!write_metadata(c.file_metadata, p, meta, npix, el, az, lists ? &counts : nullptr, lists ? &points : nullptr, max_points))
                throw std::runtime_error("cannot write " + c.file_metadata);
        }
        printf("%.3f: Done.\n", t());
        atmrt_host_free(rgb), atmrt_host_free(meta);
        atmrt_group_destroy(group);
        return 0;
    } catch (const std::exception& e) {
        atmrt_host_free(rgb), atmrt_host_free(meta);
        if (group) atmrt_group_destroy(group);
        fail(ATMRT_ERR_INVALID, e.what());
        fprintf(stderr, "ERROR: %s\n", e.what());  // main.rs:36-38
        return 1;
    }
}

// Parse-only entry for tests: YAML + CLI -> params/objects, no GPU needed.
extern "C" int atmrt_host_parse_config(int argc, const char* const* argv, atmrt_params* params, atmrt_object* objects, int max_objects,
                                       int* nobjects, char* terrain_folder, size_t folder_cap, char* output_file, size_t file_cap,
                                       char* meta_file, size_t meta_cap) {
    try {
        Config c = read_config(argc, argv);
        if (params) *params = into_params(c);
        if (nobjects) *nobjects = (int)c.objects.size();
        for (int i = 0; objects && i < max_objects && i < (int)c.objects.size(); ++i) objects[i] = c.objects[i].o;
        auto put = [](char* dst, size_t cap, const std::string& s) {
            if (dst && cap) snprintf(dst, cap, "%s", s.c_str());
        };
        put(terrain_folder, folder_cap, c.terrain_folder);
        put(output_file, file_cap, c.file);
        put(meta_file, meta_cap, c.file_metadata);
        return 0;
    } catch (const std::exception& e) {
        return fail(ATMRT_ERR_INVALID, e.what());
    }
}

// Parse-only: the overlay keys of `output` (params.rs:403-410) as read_config lowers them.
extern "C" int atmrt_host_parse_overlays(int argc, const char* const* argv, atmrt_host_tick* ticks, int max_ticks, int* nticks,
                                         atmrt_host_tick* vertical_ticks, int max_vertical_ticks, int* nvertical_ticks, int* show_eye_level,
                                         int* show_flat_horizon) {
    try {
        Config c = read_config(argc, argv);
        if (nticks) *nticks = (int)c.ticks.size();
        if (nvertical_ticks) *nvertical_ticks = (int)c.vertical_ticks.size();
        for (int i = 0; ticks && i < max_ticks && i < (int)c.ticks.size(); ++i) ticks[i] = c.ticks[(size_t)i];
        for (int i = 0; vertical_ticks && i < max_vertical_ticks && i < (int)c.vertical_ticks.size(); ++i) vertical_ticks[i] = c.vertical_ticks[(size_t)i];
        if (show_eye_level) *show_eye_level = c.show_eye_level;
        if (show_flat_horizon) *show_flat_horizon = c.show_flat_horizon;
        return 0;
    } catch (const std::exception& e) {
        return fail(ATMRT_ERR_INVALID, e.what());
    }
}

// ---------------------------------------------------------------------------------------------
// The reference's three text dumpers: the same flags, the same text layout (`{}` of an f64 is
// atmrt_fmt_f64), the numbers from the device through the C-ABI probes. `parse_config(filename)` (params.rs:678-692)
// reads the YAML only -- these subcommands take no `gen` flags.
// ---------------------------------------------------------------------------------------------
namespace {

struct DumpArgs {
    std::string input;
    std::map<std::string, std::string> val;
    std::map<std::string, bool> flag;
};

// clap's AllowLeadingHyphen: a value may start with '-' (negative numbers)
DumpArgs parse_dump_args(int argc, const char* const* argv, const std::map<std::string, std::string>& takes,
                         const std::map<std::string, std::string>& flags) {
    DumpArgs a;
    for (int i = 0; i < argc; ++i) {
        const std::string s = argv[i];
        auto t = takes.find(s);
        if (t != takes.end()) {
            if (i + 1 >= argc) throw std::runtime_error("missing value for " + s);
            a.val[t->second] = argv[++i];
            continue;
        }
        auto f = flags.find(s);
        if (f != flags.end()) {
            a.flag[f->second] = true;
            continue;
        }
        if (a.input.empty() && (s.empty() || s[0] != '-')) {
            a.input = s;
            continue;
        }
        throw std::runtime_error("unknown argument " + s);
    }
    if (a.input.empty()) throw std::runtime_error("please provide an input file");
    return a;
}

double dump_num(const DumpArgs& a, const char* key, double def, const char* what) {
    auto it = a.val.find(key);
    if (it == a.val.end()) return def;
    char* end = nullptr;
    const double v = strtod(it->second.c_str(), &end);
    if (end == it->second.c_str() || *end != '\0') throw std::runtime_error(std::string("please provide a valid ") + what);
    return v;
}

Config parse_config_file(const std::string& filename) {  // params.rs:678-692
    Config c;
    c.atmosphere = us_76();
    apply_yaml(parse_yaml(read_file(filename)), &c);
    return c;
}

struct Ctx {  // a context that is destroyed on every exit path
    atmrt_ctx* h = nullptr;
    ~Ctx() {
        if (h) atmrt_destroy(h);
    }
    void check(int rc, const char* what) const {
        if (rc != 0) throw std::runtime_error(std::string(what) + ": " + atmrt_last_error(h));
    }
};

int dump_fail(const std::exception& e) {
    fail(ATMRT_ERR_INVALID, e.what());
    fprintf(stderr, "ERROR: %s\n", e.what());  // main.rs:36-38
    return 1;
}

}  // namespace

// output-ray-paths (ray_path.rs:6-106): h(x) of refracted rays cast at a range of elevation angles.
extern "C" int atmrt_host_output_ray_paths(int argc, const char* const* argv) {
    try {
        const DumpArgs a = parse_dump_args(argc, argv,
                                           {{"-h", "height"}, {"--height", "height"}, {"-a", "min_angle"}, {"--min-ang", "min_angle"},
                                            {"-b", "max_angle"}, {"--max-ang", "max_angle"}, {"-s", "angle_step"}, {"--angle-step", "angle_step"},
                                            {"-r", "ray_step"}, {"--ray-step", "ray_step"}, {"-c", "cutoff_dist"}, {"--cutoff-dist", "cutoff_dist"},
                                            {"-o", "output_step"}, {"--output-step", "output_step"}},
                                           {});
        const double height = dump_num(a, "height", 2.0, "observer height");
        const double min_ang = dump_num(a, "min_angle", -1.0, "minimum altitude"), max_ang = dump_num(a, "max_angle", 1.0, "maximum altitude");
        const double step = dump_num(a, "angle_step", 0.1, "step size"), ray_step = dump_num(a, "ray_step", 50.0, "ray step size");
        const double cutoff = dump_num(a, "cutoff_dist", 10000.0, "cutoff distance"), output_step = dump_num(a, "output_step", 50.0, "output step");
        if (!(step > 0.0)) throw std::runtime_error("step must be positive");
        if (!(ray_step > 0.0)) throw std::runtime_error("ray step must be positive");
        const Config c = parse_config_file(a.input);
        const atmrt_params p = into_params(c);
        // the angles: `ang += step` from min_ang while ang <= max_ang (ray_path.rs:65-94)
        std::vector<double> angles;
        for (double ang = min_ang; ang <= max_ang; ang += step) {
            fprintf(stderr, "Elevation angle %s (min=%s, max=%s)\n", atmrt_fmt_f64(ang).c_str(), atmrt_fmt_f64(min_ang).c_str(), atmrt_fmt_f64(max_ang).c_str());
            angles.push_back(ang);
        }
        Ctx ctx;
        ctx.check(atmrt_create(0, &ctx.h), "atmrt_create");
        ctx.check(atmrt_set_params(ctx.h, &p), "atmrt_set_params");
        // every ray is stepped until x >= cutoff (ray_path.rs:88-90); x is the same running sum for every ray
        int nsteps = 0;
        std::vector<double> x;
        for (int cap = 1024;; cap *= 2) {
            x.assign((size_t)cap, 0.0);
            const double none = 0.0;  // (no rays in this call: only RayState::x is wanted)
            ctx.check(atmrt_ray_paths(ctx.h, height, &none, 0, ray_step, cap, x.data(), nullptr), "atmrt_ray_paths");
            int k = 0;
            while (k < cap && !(x[(size_t)k] >= cutoff)) ++k;
            if (k < cap) {
                nsteps = k + 1;
                break;
            }
            if (cap > (1 << 26)) throw std::runtime_error("cutoff distance / ray step too large");
        }
        x.resize((size_t)nsteps);
        std::vector<double> h((size_t)angles.size() * nsteps);
        if (!angles.empty())
            ctx.check(atmrt_ray_paths(ctx.h, height, angles.data(), (int)angles.size(), ray_step, nsteps, x.data(), h.data()), "atmrt_ray_paths");
        // a state is output when an output_step boundary falls within ray_step / 2 of it (ray_path.rs:80-87)
        std::vector<int> keep;
        for (int k = 0; k < nsteps; ++k)
            if (std::floor((x[(size_t)k] - ray_step / 2.0) / output_step) != std::floor((x[(size_t)k] + ray_step / 2.0) / output_step)) keep.push_back(k);
        std::string out;
        for (size_t i = 0; i <= keep.size(); ++i) {
            out += atmrt_fmt_f64(i == 0 ? 0.0 : x[(size_t)keep[i - 1]]) + "\t";
            for (size_t r = 0; r < angles.size(); ++r) out += atmrt_fmt_f64(i == 0 ? height : h[r * nsteps + (size_t)keep[i - 1]]) + "\t";
            out += "\n";
        }
        fwrite(out.data(), 1, out.size(), stdout);
        return 0;
    } catch (const std::exception& e) {
        return dump_fail(e);
    }
}

// output-elev-profile (elev_profile.rs:9-67): terrain elevation against distance along one azimuth from the observer.
extern "C" int atmrt_host_output_elev_profile(int argc, const char* const* argv) {
    try {
        const DumpArgs a = parse_dump_args(argc, argv,
                                           {{"-a", "azim"}, {"--azim", "azim"}, {"-s", "step"}, {"--step", "step"}, {"-c", "cutoff_dist"}, {"--cutoff-dist", "cutoff_dist"}}, {});
        const double azim = dump_num(a, "azim", 0.0, "azimuth"), step = dump_num(a, "step", 50.0, "step size");
        const double cutoff = dump_num(a, "cutoff_dist", 10000.0, "cutoff distance");
        if (!(step > 0.0)) throw std::runtime_error("step must be positive");
        const Config c = parse_config_file(a.input);
        const HostTerrain terrain = load_terrain_folder(c.terrain_folder);
        const atmrt_params p = into_params(c);
        std::vector<double> xs;
        for (double x = 0.0; x <= cutoff; x += step) xs.push_back(x);  // elev_profile.rs:53-60
        Ctx ctx;
        ctx.check(atmrt_create(0, &ctx.h), "atmrt_create");
        ctx.check(atmrt_set_terrain(ctx.h, terrain.descs.data(), (int)terrain.descs.size(), terrain.ptrs().data()), "atmrt_set_terrain");
        ctx.check(atmrt_set_params(ctx.h, &p), "atmrt_set_params");
        std::vector<double> elev(xs.size());
        ctx.check(atmrt_elev_profile(ctx.h, azim, xs.data(), (int)xs.size(), nullptr, nullptr, elev.data()), "atmrt_elev_profile");
        std::string out;
        for (size_t i = 0; i < xs.size(); ++i) out += atmrt_fmt_f64(xs[i]) + "\t" + atmrt_fmt_f64(elev[i]) + "\n";
        fwrite(out.data(), 1, out.size(), stdout);
        return 0;
    } catch (const std::exception& e) {
        return dump_fail(e);
    }
}

// output-atm (atm_printer.rs:6-49): temperature, pressure and relative humidity against altitude.
extern "C" int atmrt_host_output_atm(int argc, const char* const* argv) {
    try {
        const DumpArgs a = parse_dump_args(argc, argv,
                                           {{"-a", "min_altitude"}, {"--min-alt", "min_altitude"}, {"-b", "max_altitude"}, {"--max-alt", "max_altitude"}, {"-s", "step"}, {"--step", "step"}},
                                           {{"-c", "celsius"}, {"--celsius", "celsius"}});
        const double min_alt = dump_num(a, "min_altitude", 0.0, "minimum altitude"), max_alt = dump_num(a, "max_altitude", 1000.0, "maximum altitude");
        const double step = dump_num(a, "step", 0.2, "step size");
        const bool celsius = a.flag.count("celsius") != 0;
        if (!(step > 0.0)) throw std::runtime_error("step must be positive");  // (the reference would loop forever)
        const Config c = parse_config_file(a.input);
        const atmrt_params p = into_params(c);
        std::vector<double> alts;
        for (double alt = min_alt; alt <= max_alt; alt += step) alts.push_back(alt);  // atm_printer.rs:35-46
        Ctx ctx;
        ctx.check(atmrt_create(0, &ctx.h), "atmrt_create");
        ctx.check(atmrt_set_params(ctx.h, &p), "atmrt_set_params");
        std::vector<double> t(alts.size()), pr(alts.size());
        ctx.check(atmrt_atmosphere_probe(ctx.h, alts.data(), (int)alts.size(), t.data(), pr.data(), nullptr), "atmrt_atmosphere_probe");
        std::string out;
        for (size_t i = 0; i < alts.size(); ++i)
            out += atmrt_fmt_f64(alts[i]) + " " + atmrt_fmt_f64(t[i] - (celsius ? 273.15 : 0.0)) + " " + atmrt_fmt_f64(pr[i]) + " " +
                   atmrt_fmt_f64(c.atmosphere.humidity) + "\n";
        fwrite(out.data(), 1, out.size(), stdout);
        return 0;
    } catch (const std::exception& e) {
        return dump_fail(e);
    }
}
