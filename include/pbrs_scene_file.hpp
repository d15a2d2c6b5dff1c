// pbrs_scene_file.hpp -- pbrt-v3-subset scene files -> pbrs::Scene (C++17, header only).
//
// The input side of the boundary, as the reference reads it:
//   lexer   scene_parser/src/token.rs (token set; no exponent floats, :112-114; `#` comments),
//           scene_parser/src/lexer.rs:41-56 (`Include`)
//   parser  scene_parser/src/parser.rs (scene-wide options, WorldBegin..WorldEnd, attribute and
//           transform blocks; a one-element bracket list is a Number, :230-239)
//   loader  scene/src/loader.rs (camera :91-135; world items :164-305; sphere / trianglemesh /
//           plymesh :307-389 with scene/src/plyloader.rs; area lights on spheres and PLY meshes :396-434; distant / point / infinite lights; the
//           materials glass / mirror / matte / metal / plastic / uber / substrate :483-714;
//           `Rotate` with the negated angle :792-798; quirks Q16 of SURVEY.md)
// Where the reference panics or hits unimplemented!() this throws pbrs::Error(PBRS_ERR_UNSUPPORTED).
// pbrs_b200/pbrt_loader.py is the same loader in Python (tests compare the two).
#pragma once
#include <cctype>
#include <fstream>
#include <functional>
#include <sstream>

#include "pbrs_gpu.hpp"

namespace pbrs {
namespace scene_file {

struct Token {
    enum Kind { Kw, Num, Str, LB, RB, End } kind;
    std::string text;
    float num = 0.0f;
};
inline Error unsupported(const std::string &m) { return Error(PBRS_ERR_UNSUPPORTED, m); }

inline std::string dir_of(const std::string &path) {
    auto p = path.find_last_of('/');
    return p == std::string::npos ? "." : path.substr(0, p);
}
inline std::string read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(PBRS_ERR_INVALID_ARG, "open file failed: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

inline void tokenize(const std::string &text, const std::string &root, std::vector<Token> &out) {
    size_t i = 0, n = text.size();
    while (i < n) {
        char c = text[i];
        if (c == ' ' || c == '\t' || c == '\n' || c == '\f' || c == '\r') { ++i; continue; }
        if (c == '#') { while (i < n && text[i] != '\n') ++i; continue; }
        if (c == '[') { out.push_back({Token::LB, "["}); ++i; continue; }
        if (c == ']') { out.push_back({Token::RB, "]"}); ++i; continue; }
        if (c == '"') {
            size_t j = text.find_first_of("\"\n", i + 1);
            if (j == std::string::npos || text[j] != '"' || j == i + 1) throw Error(PBRS_ERR_INVALID_ARG, "lexer: bad string literal");
            out.push_back({Token::Str, text.substr(i + 1, j - i - 1)});
            i = j + 1;
            continue;
        }
        if (std::isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') {
            size_t j = i;
            if (text[j] == '-' || text[j] == '+') ++j;
            size_t d0 = j;
            while (j < n && std::isdigit((unsigned char)text[j])) ++j;
            bool had_int = j > d0;
            if (j < n && text[j] == '.') { ++j; size_t f0 = j; while (j < n && std::isdigit((unsigned char)text[j])) ++j; if (!had_int && j == f0) throw Error(PBRS_ERR_INVALID_ARG, "lexer: bad number"); }
            else if (!had_int) throw Error(PBRS_ERR_INVALID_ARG, "lexer: bad number");
            Token t{Token::Num, text.substr(i, j - i)};
            t.num = std::strtof(t.text.c_str(), nullptr);
            out.push_back(t);
            i = j;
            continue;
        }
        if (std::isalpha((unsigned char)c)) {
            size_t j = i;
            while (j < n && std::isalpha((unsigned char)text[j])) ++j;
            std::string kw = text.substr(i, j - i);
            i = j;
            if (kw == "Include") {  // lexer.rs:41-56
                std::vector<Token> inc;
                size_t k = i;
                while (k < n && std::isspace((unsigned char)text[k])) ++k;
                if (k >= n || text[k] != '"') throw Error(PBRS_ERR_INVALID_ARG, "should have a file name after Include");
                size_t e = text.find('"', k + 1);
                std::string path = root + "/" + text.substr(k + 1, e - k - 1);
                tokenize(read_file(path), dir_of(path), out);
                i = e + 1;
                continue;
            }
            static const char *known[] = {"LookAt", "Camera", "Integrator", "Accelerator", "Sampler", "Film", "PixelFilter", "Filter", "WorldBegin", "WorldEnd",
                "AttributeBegin", "AttributeEnd", "TransformBegin", "TransformEnd", "LightSource", "AreaLightSource", "Material", "Shape", "Texture", "Identity",
                "Translate", "Scale", "Rotate", "CoordinateSystem", "CoordSysTransform", "Transform", "ConcatTransform", "ReverseOrientation", "MediumInterface",
                "NamedMedium", "MakeNamedMedium", "NamedMaterial", "MakeNamedMaterial", "ObjectBegin", "ObjectEnd", "ObjectInstance"};
            bool ok = false;
            for (const char *k : known) ok = ok || kw == k;
            if (!ok) throw Error(PBRS_ERR_INVALID_ARG, "lexer: unknown directive " + kw);
            out.push_back({Token::Kw, kw});
            continue;
        }
        throw Error(PBRS_ERR_INVALID_ARG, std::string("lexer error at '") + c + "'");
    }
}

// scene_parser/src/ast.rs ArgValue / ParameterSet
// ---------------------------------------------------------------------------------------------
// PLY meshes: scene/src/plyloader.rs:69-256 (binary payloads, float vertex properties, polygons
// fanned) + geometry/src/lib.rs:16-32 (compute_normals).  Upstream the file is cut off right after
// the normals are computed; the tail here is what the signature and call sites require: uv
// defaults to (0, 0) and the arrays become TriangleMeshRaw.  pbrt_loader.load_ply is the same code.
// ---------------------------------------------------------------------------------------------
struct PlyMesh {
    std::vector<float> P, N, UV;
    std::vector<uint32_t> idx;
};
inline std::vector<float> compute_normals(const std::vector<float> &P, const std::vector<uint32_t> &idx) {
    std::vector<float> acc(P.size(), 0.0f);
    for (size_t t = 0; t + 2 < idx.size(); t += 3) {
        const float *p0 = &P[3 * idx[t]], *p1 = &P[3 * idx[t + 1]], *p2 = &P[3 * idx[t + 2]];
        const float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        for (int v = 0; v < 3; ++v)
            for (int c = 0; c < 3; ++c) acc[3 * idx[t + v] + c] += n[c];
    }
    for (size_t v = 0; v + 2 < acc.size(); v += 3) {
        float n2 = (acc[v] * acc[v] + acc[v + 1] * acc[v + 1]) + acc[v + 2] * acc[v + 2];
        if (!(n2 != 0.0f && std::isfinite(n2))) throw Error(PBRS_ERR_UNSUPPORTED, "compute_normals: a vertex has no (finite, non-zero) normal; Vec3::hat panics upstream");
        float inv = 1.0f / std::sqrt(n2);
        for (int c = 0; c < 3; ++c) acc[v + c] = acc[v + c] * inv;
    }
    return acc;
}
inline PlyMesh load_ply(const std::string &path) {
    const std::string data = read_file(path);
    size_t pos = 0;
    auto line = [&]() {
        size_t end = data.find('\n', pos);
        if (end == std::string::npos) throw Error(PBRS_ERR_INVALID_ARG, "ply: unexpected end of header");
        std::string out = data.substr(pos, end - pos);
        pos = end + 1;
        while (!out.empty() && (out.back() == '\r' || out.back() == ' ')) out.pop_back();
        size_t b = out.find_first_not_of(' ');
        return b == std::string::npos ? std::string() : out.substr(b);
    };
    auto split = [](const std::string &s) {
        std::vector<std::string> w;
        std::stringstream ss(s);
        std::string part;
        while (std::getline(ss, part, ' ')) w.push_back(part);
        return w;
    };
    auto type_size = [](const std::string &n) { return (n == "uchar" || n == "uint8") ? 1 : n == "short" ? 2 : (n == "int" || n == "uint") ? 4 : 0; };  // :14-21
    auto count = [](const std::string &n, long &out) {
        // at most 10 digits and below 2^31: longer counts would overflow the size arithmetic below (and std::stol throws)
        if (n.empty() || n.size() > 10 || n.find_first_not_of("0123456789") != std::string::npos) return false;
        out = std::stol(n);
        return out < (1L << 31);
    };
    if (line() != "ply") throw Error(PBRS_ERR_INVALID_ARG, "ply: Header isn't ply");
    std::vector<std::string> fw = split(line());
    if (fw.size() < 2 || fw[0] != "format") throw Error(PBRS_ERR_INVALID_ARG, "ply: Format line is bad");
    const std::string fmt = fw[1];
    if (fmt != "ascii" && fmt != "binary_little_endian" && fmt != "binary_big_endian") throw Error(PBRS_ERR_INVALID_ARG, "ply: Unrecognized format string: " + fmt);
    std::vector<std::string> props;
    long nv = -1, nf = -1;
    int len_size = 0, el_size = 0;
    while (true) {
        std::string ln = line();
        if (ln == "end_header") break;
        if (ln.rfind("comment", 0) == 0) continue;
        std::vector<std::string> w = split(ln);
        if (w.size() < 3) throw Error(PBRS_ERR_INVALID_ARG, "ply: Can't handle the line " + ln);
        if (w.size() == 3 && w[0] == "element" && w[1] == "vertex") { if (!count(w[2], nv)) nv = -1; }
        else if (w.size() == 3 && w[0] == "element" && w[1] == "face") { if (!count(w[2], nf)) nf = -1; }
        else if (w.size() == 3 && w[0] == "property" && w[1] == "float") props.push_back(w[2]);
        else if (w.size() == 5 && w[0] == "property" && w[1] == "list" && w[4] == "vertex_indices") { len_size = type_size(w[2]); el_size = type_size(w[3]); }
        else if (w[0] == "property") throw unsupported("ply: unsupported property line '" + ln + "' (only float vertex properties and a vertex_indices list)");
    }
    if (nv < 0 || nf < 0 || !len_size || !el_size) throw Error(PBRS_ERR_INVALID_ARG, "ply: header lacks vertex / face counts or the vertex_indices list");
    if (fmt == "ascii") throw unsupported("ply: ascii payloads are not supported (bytes_to_f32 panics upstream)");
    const bool le = fmt == "binary_little_endian";
    auto u_at = [&](size_t at, int size) {
        uint32_t v = 0;
        for (int b = 0; b < size; ++b) v |= uint32_t((unsigned char)data[at + (le ? b : size - 1 - b)]) << (8 * b);
        return v;
    };
    const size_t stride = props.size();
    if (pos > data.size() || (stride != 0 && size_t(nv) > (data.size() - pos) / (stride * 4))) throw Error(PBRS_ERR_INVALID_ARG, "ply: vertex block is truncated");
    std::vector<float> vb(size_t(nv) * stride);
    for (size_t k = 0; k < vb.size(); ++k) { uint32_t u = u_at(pos + 4 * k, 4); std::memcpy(&vb[k], &u, 4); }
    pos += vb.size() * 4;
    PlyMesh m;
    for (long f = 0; f < nf; ++f) {
        if (pos + len_size > data.size()) throw Error(PBRS_ERR_INVALID_ARG, "ply: face block is truncated");
        uint32_t n = u_at(pos, len_size);
        pos += len_size;
        if (pos + size_t(n) * el_size > data.size()) throw Error(PBRS_ERR_INVALID_ARG, "ply: face block is truncated");
        if (n == 0) throw Error(PBRS_ERR_INVALID_ARG, "ply: empty face (list_length - 1 underflows upstream)");
        std::vector<uint32_t> face(n);
        for (uint32_t k = 0; k < n; ++k) face[k] = u_at(pos + size_t(k) * el_size, el_size);
        pos += size_t(n) * el_size;
        if (n == 3) m.idx.insert(m.idx.end(), {face[0], face[1], face[2]});
        else for (uint32_t k = 1; k + 1 < n; ++k) m.idx.insert(m.idx.end(), {face[0], face[k], face[k + 1]});  // fan, :184-190
    }
    for (uint32_t v : m.idx) if (v >= uint32_t(nv)) throw Error(PBRS_ERR_INVALID_ARG, "ply: vertex index out of range");
    int ox = -1, oy = -1, oz = -1, onx = -1, ony = -1, onz = -1, ou = -1, ov = -1;
    for (size_t k = 0; k < stride; ++k) {
        const std::string &n = props[k];
        if (n == "x") ox = int(k); else if (n == "y") oy = int(k); else if (n == "z") oz = int(k);
        else if (n == "nx") onx = int(k); else if (n == "ny") ony = int(k); else if (n == "nz") onz = int(k);
        else if (n == "u") ou = int(k); else if (n == "v") ov = int(k);
    }
    if (ox < 0 || oy < 0 || oz < 0) throw Error(PBRS_ERR_INVALID_ARG, "ply: position xyz: some missing");
    const bool has_n = onx >= 0 && ony >= 0 && onz >= 0, has_uv = ou >= 0 && ov >= 0;
    for (long v = 0; v < nv; ++v) {
        const float *row = &vb[size_t(v) * stride];
        m.P.insert(m.P.end(), {row[ox], row[oy], row[oz]});
        if (has_n) m.N.insert(m.N.end(), {row[onx], row[ony], row[onz]});
        if (has_uv) { m.UV.push_back(row[ou]); m.UV.push_back(row[ov]); } else { m.UV.push_back(0.0f); m.UV.push_back(0.0f); }
    }
    if (!has_n) m.N = compute_normals(m.P, m.idx);
    return m;
}

struct Arg {
    enum Kind { Number, Numbers, String } kind = Number;
    float number = 0.0f;
    std::vector<float> numbers;
    std::string str;
};
struct Params {
    std::vector<std::pair<std::string, Arg>> kv;  // insertion order; keys unique
    bool extract(const std::string &key, Arg &out) {
        for (size_t i = 0; i < kv.size(); ++i) if (kv[i].first == key) { out = kv[i].second; kv.erase(kv.begin() + i); return true; }
        return false;
    }
    // ast.rs:58-69: a key one of whose space-separated parts equals `pattern`
    bool extract_substr(const std::string &pattern, std::string &key, Arg &out) {
        for (size_t i = 0; i < kv.size(); ++i) {
            std::stringstream ss(kv[i].first);
            std::string part;
            while (std::getline(ss, part, ' ')) if (part == pattern) { key = kv[i].first; out = kv[i].second; kv.erase(kv.begin() + i); return true; }
        }
        return false;
    }
    bool lookup_f32(const std::string &key, float &out) const {
        for (auto &p : kv) if (p.first == key && p.second.kind == Arg::Number) { out = p.second.number; return true; }
        return false;
    }
};

struct Loader {
    std::string root;
    std::vector<Token> t;
    size_t i = 0;
    // loader state (scene/src/loader.rs:22-37)
    std::vector<InstanceTransform> ctm{InstanceTransform::identity()};
    MaterialRef cur_mtl;
    bool has_area_l = false;
    Color area_l;
    std::map<std::string, TextureRef> named_tex;
    std::map<std::string, MaterialRef> named_mtl;
    std::vector<Instance> instances;
    std::vector<light::DeltaLight> delta;
    std::vector<light::DiffuseAreaLight> area;
    int env_kind = 0;
    Color env_color = Color::black(), env_scale = Color::white();
    TextureRef env_image;
    // the image decoder is the caller's business (the reference uses the `png` crate): path -> (w, h, rgb8)
    std::function<TextureRef(const std::string &)> load_image;

    const Token &peek() const { return t[i]; }
    bool is_kw(const char *k) const { return peek().kind == Token::Kw && peek().text == k; }
    void expect_kw(const char *k) { if (!is_kw(k)) throw Error(PBRS_ERR_INVALID_ARG, std::string("expected ") + k + ", found " + peek().text); ++i; }
    std::string quoted() { if (peek().kind != Token::Str) throw Error(PBRS_ERR_INVALID_ARG, "expected a quoted string, found " + peek().text); return t[i++].text; }
    std::vector<float> numbers() { std::vector<float> v; while (peek().kind == Token::Num) v.push_back(t[i++].num); return v; }

    Params parameter_list() {  // parser.rs:216-257
        Params ps;
        while (peek().kind == Token::Str) {
            std::string key = quoted();
            Arg a;
            if (peek().kind == Token::LB) {
                ++i;
                if (peek().kind == Token::Num) {
                    std::vector<float> v = numbers();
                    if (v.size() == 1) { a.kind = Arg::Number; a.number = v[0]; } else { a.kind = Arg::Numbers; a.numbers = v; }
                } else if (peek().kind == Token::Str) { a.kind = Arg::String; a.str = quoted(); }
                else throw Error(PBRS_ERR_INVALID_ARG, "only numbers or quoted strings allowed");
                if (peek().kind != Token::RB) throw Error(PBRS_ERR_INVALID_ARG, "expected ]");
                ++i;
            } else if (peek().kind == Token::Str) { a.kind = Arg::String; a.str = quoted(); }
            else if (peek().kind == Token::Num) { a.kind = Arg::Number; a.number = t[i++].num; }
            else throw Error(PBRS_ERR_INVALID_ARG, "unexpected token after key");
            bool replaced = false;
            for (auto &p : ps.kv) if (p.first == key) { p.second = a; replaced = true; }
            if (!replaced) ps.kv.emplace_back(key, a);
        }
        return ps;
    }

    static bool starts_transform(const Token &k) {
        if (k.kind != Token::Kw) return false;
        for (const char *n : {"Identity", "Translate", "Scale", "Rotate", "LookAt", "Transform", "ConcatTransform", "CoordSysTransform", "CoordinateSystem"}) if (k.text == n) return true;
        return false;
    }
    static bool starts_world_item(const Token &k) {
        if (starts_transform(k)) return true;
        if (k.kind != Token::Kw) return false;
        for (const char *n : {"Shape", "Material", "LightSource", "AreaLightSource", "Texture", "MakeNamedMaterial", "ObjectInstance", "AttributeBegin", "ObjectBegin",
                              "TransformBegin", "NamedMaterial", "ReverseOrientation"}) if (k.text == n) return true;
        return false;
    }

    struct Xf { std::string kind; std::vector<float> n; };
    Xf parse_transform() {  // parser.rs:259-311
        Xf x{t[i++].text, {}};
        if (x.kind == "Identity") return x;
        x.n = numbers();
        if (x.kind == "Translate" && x.n.size() != 3) throw Error(PBRS_ERR_INVALID_ARG, "wrong number of numbers after translation");
        if (x.kind == "Scale" && x.n.size() < 3) throw Error(PBRS_ERR_INVALID_ARG, "Scale needs 3 numbers");
        if (x.kind == "Rotate" && x.n.size() != 4) throw Error(PBRS_ERR_INVALID_ARG, "4 numbers expected in Rotate");
        if (x.kind == "LookAt" && x.n.size() != 9) throw Error(PBRS_ERR_INVALID_ARG, "wrong numbers of floats in LookAt");
        if (x.kind != "Translate" && x.kind != "Scale" && x.kind != "Rotate" && x.kind != "LookAt") throw unsupported("transform directive " + x.kind + " (unimplemented!() upstream)");
        return x;
    }
    static InstanceTransform to_affine(const Xf &x) {  // loader.rs:784-803
        if (x.kind == "Identity") return InstanceTransform::identity();
        if (x.kind == "Translate") return InstanceTransform::translater({x.n[0], x.n[1], x.n[2]});
        if (x.kind == "Scale") return InstanceTransform::scaler({x.n[0], x.n[1], x.n[2]});
        if (x.kind == "Rotate") return InstanceTransform::rotater({x.n[1], x.n[2], x.n[3]}, Angle::new_rad(-Angle::new_deg(x.n[0]).radian));  // negated, :792-798
        throw unsupported("unsupported lookat in modeling step");
    }

    // ---- colours / textures / numbers ----
    static Color constant_color(const std::string &key, const std::vector<float> &n) {  // :758-766
        std::string type = key.substr(0, key.find(' '));
        if ((type == "rgb" || type == "color") && n.size() >= 3) return {n[0], n[1], n[2]};
        if (type == "xyz" && n.size() >= 3) {  // Color::from_xyz, radiometry/src/color.rs:30-36
            const float x = n[0], y = n[1], z = n[2];
            return {3.240479f * x - 1.537150f * y - 0.498535f * z, -0.969256f * x + 1.875991f * y + 0.041556f * z,
                    0.055648f * x - 0.204043f * y + 1.057311f * z};
        }
        throw unsupported("colour type '" + type + "' (blackbody / spectrum are load-time spectra over the CIE tables: out of scope)");
    }
    Color color_arg(Params &ps, const char *name, Color dflt) {
        std::string key; Arg a;
        if (!ps.extract_substr(name, key, a)) return dflt;
        if (a.kind == Arg::Numbers) return constant_color(key, a.numbers);
        if (a.kind == Arg::Number) return Color::gray(a.number);
        throw unsupported(std::string("textured ") + name + " (unimplemented!() upstream)");
    }
    TextureRef tex_arg(Params &ps, const char *name, float dflt_gray, bool optional = false) {  // solid_or_image_tex, :737-752
        std::string key; Arg a;
        if (!ps.extract_substr(name, key, a)) return optional ? nullptr : tex::Solid::create(Color::gray(dflt_gray));
        if (a.kind == Arg::Numbers) return tex::Solid::create(constant_color(key, a.numbers));
        if (a.kind == Arg::Number) return tex::Solid::create(Color::gray(a.number));
        auto it = named_tex.find(a.str);
        if (it == named_tex.end()) throw Error(PBRS_ERR_INVALID_ARG, "unknown texture " + a.str);
        return it->second;
    }
    static float num_arg(Params &ps, const char *name, float dflt) {
        std::string key; Arg a;
        if (!ps.extract_substr(name, key, a)) return dflt;
        if (a.kind != Arg::Number) throw Error(PBRS_ERR_INVALID_ARG, std::string(name) + " value isn't a number");
        return a.number;
    }
    static bool bool_arg(Params &ps, const char *name, bool dflt) {
        std::string key; Arg a;
        if (!ps.extract_substr(name, key, a)) return dflt;
        if (a.kind != Arg::String || (a.str != "true" && a.str != "false")) throw Error(PBRS_ERR_INVALID_ARG, "invalid boolean string");
        return a.str == "true";
    }

    MaterialRef material(const std::string &impl, Params &ps) {  // :483-714
        const Color copper{0.19547f, 0.925682f, 1.102186f};  // preset::copper_fresnel().0
        if (impl == "glass") { Color kr = color_arg(ps, "Kr", Color::white()), kt = color_arg(ps, "Kt", Color::white()); return mtl::Dielectric::create(num_arg(ps, "eta", 1.5f), kr, kt); }
        if (impl == "mirror") return mtl::Mirror::create(color_arg(ps, "Kr", Color::gray(0.9f)));
        if (impl == "matte") { TextureRef kd = tex_arg(ps, "Kd", 0.5f); Arg a; ps.extract("sigma", a); return mtl::Lambertian::textured(kd); }
        if (impl == "metal") {
            float rough = num_arg(ps, "roughness", 0.01f); Arg a; ps.extract("remaproughness", a);
            Color eta = color_arg(ps, "eta", copper), k = color_arg(ps, "k", copper);  // Q16: k defaults to copper ETA (:560)
            return mtl::Metal::from_ior(eta, k, rough);
        }
        if (impl == "plastic") {
            Color kd = color_arg(ps, "Kd", Color::gray(0.25f)), ks = color_arg(ps, "Ks", Color::gray(0.25f));
            float rough = num_arg(ps, "roughness", 0.1f);
            return mtl::Plastic::create(kd, ks, rough, bool_arg(ps, "remaproughness", true));
        }
        if (impl == "uber") {
            TextureRef kd = tex_arg(ps, "Kd", 0.25f), ks = tex_arg(ps, "Ks", 0.25f), kr = tex_arg(ps, "Kr", 0, true), kt = tex_arg(ps, "Kt", 0, true);
            float ur = num_arg(ps, "uroughness", 0.0f), vr = num_arg(ps, "vroughness", 0.0f), r = num_arg(ps, "roughness", 0.0f), eta = num_arg(ps, "eta", 1.5f);
            bool remap = bool_arg(ps, "remaproughness", true);
            // Q16: `opacity` re-reads "eta" (already extracted) and is therefore always 1 (:644)
            return mtl::Uber::create(kd, ks, kr, kt, ur == vr ? r : ur, ur == vr ? r : vr, eta, 1.0f, remap);
        }
        if (impl == "substrate") { TextureRef kd = tex_arg(ps, "Kd", 0.5f), ks = tex_arg(ps, "Ks", 0.5f); return mtl::Substrate::create(kd, ks); }
        throw unsupported("not recognized material: " + impl);
    }

    ShapeRef shape(const std::string &impl, Params &ps) {  // :307-389
        if (impl == "sphere") { float r = 1.0f; ps.lookup_f32("float radius", r); return shape::Sphere::create({0, 0, 0}, r); }
        if (impl == "trianglemesh") {
            Arg P, uv, idx, nrm; std::string key;
            if (!ps.extract("point P", P) || P.kind != Arg::Numbers) throw Error(PBRS_ERR_INVALID_ARG, "missing points");
            std::vector<float> uvs;
            if (ps.extract("float uv", uv) || ps.extract("float st", uv)) uvs = uv.numbers;
            if (!ps.extract("integer indices", idx) || idx.kind != Arg::Numbers) throw Error(PBRS_ERR_INVALID_ARG, "missing indices");
            std::vector<uint32_t> tri(idx.numbers.size());
            for (size_t k = 0; k < tri.size(); ++k) tri[k] = (uint32_t)idx.numbers[k];
            std::vector<float> normals;
            if (ps.extract_substr("normal", key, nrm)) normals = nrm.numbers;
            return shape::TriangleMesh::from_soa(P.numbers, normals, uvs, tri);
        }
        if (impl == "plymesh") {  // :314-331
            PlyMesh m = ply(ps);
            return shape::TriangleMesh::from_soa(m.P, m.N, m.UV, m.idx);
        }
        throw unsupported("shape of " + impl + " (loopsubdiv: pre-process)");
    }
    PlyMesh ply(Params &ps) {
        for (auto &p : ps.kv) if (p.first == "string filename" && p.second.kind == Arg::String) return load_ply(root + "/" + p.second.str);  // lookup_string
        throw Error(PBRS_ERR_INVALID_ARG, "no ply file specified");
    }

    static void apply(const float m[4][4], const float v[4], float out[3]) {
        for (int r = 0; r < 3; ++r) out[r] = ((m[0][r] * v[0] + m[1][r] * v[1]) + m[2][r] * v[2]) + m[3][r] * v[3];
    }

    void world_item() {  // loader.rs:164-305 over parser.rs:38-158
        if (starts_transform(peek())) { ctm.back() = ctm.back() * to_affine(parse_transform()); return; }
        std::string k = t[i++].text;
        if (k == "Shape") {
            std::string impl = quoted();
            Params ps = parameter_list();
            Arg a; ps.extract("alpha", a);
            const InstanceTransform &c = ctm.back();
            if (has_area_l && impl == "plymesh") {
                // parse_samplable_shape, :408-433: one IsolatedTriangle instance and one triangle light
                // (world-space vertices) per face
                PlyMesh m = ply(ps);
                MaterialRef light_mtl = mtl::DiffuseLight::create(area_l);
                for (size_t f = 0; f + 2 < m.idx.size(); f += 3) {
                    Point3 w[3], o3[3];
                    for (int v = 0; v < 3; ++v) {
                        const float *pv = &m.P[3 * m.idx[f + v]];
                        o3[v] = {pv[0], pv[1], pv[2]};
                        const float hv[4] = {pv[0], pv[1], pv[2], 1.0f};
                        float o[3];
                        apply(c.fwd, hv, o);  // SamplableShape::transformed_by, light/src/sample_shape.rs:63-68
                        w[v] = {o[0], o[1], o[2]};
                    }
                    area.emplace_back(area_l, light::SamplableShape::Triangle(w[0], w[1], w[2]));
                    instances.push_back(Instance(shape::IsolatedTriangle::create(o3[0], o3[1], o3[2]), light_mtl).with_transform(c));
                }
            } else if (has_area_l) {
                if (impl != "sphere") throw unsupported("samplable shape: " + impl);
                float r = 1.0f; ps.lookup_f32("float radius", r);
                // SamplableShape::transformed_by, light/src/sample_shape.rs:46-82
                const float X[4] = {1, 0, 0, 0}, Y[4] = {0, 1, 0, 0}, Z[4] = {0, 0, 1, 0}, O[4] = {0, 0, 0, 1};
                float tx[3], ty[3], tz[3], ce[3];
                apply(c.fwd, X, tx); apply(c.fwd, Y, ty); apply(c.fwd, Z, tz); apply(c.fwd, O, ce);
                float cr[3] = {tx[1] * ty[2] - tx[2] * ty[1], tx[2] * ty[0] - tx[0] * ty[2], tx[0] * ty[1] - tx[1] * ty[0]};
                float scale = std::cbrt(cr[0] * tz[0] + cr[1] * tz[1] + cr[2] * tz[2]);
                if (!(scale > 0.0f)) throw Error(PBRS_ERR_INVALID_ARG, "area-light transform must have positive uniform scale");
                area.emplace_back(area_l, light::SamplableShape::Sphere({ce[0], ce[1], ce[2]}, r * scale));
                instances.push_back(Instance(shape::Sphere::create({0, 0, 0}, r), mtl::DiffuseLight::create(area_l)).with_transform(c));
            } else if (cur_mtl) {
                instances.push_back(Instance(shape(impl, ps), cur_mtl).with_transform(c));
            }  // else "Neither arealight luminance or material are set": dropped (:196)
        } else if (k == "Material") { std::string impl = quoted(); Params ps = parameter_list(); cur_mtl = material(impl, ps); }
        else if (k == "LightSource") { std::string impl = quoted(); Params ps = parameter_list(); light_source(impl, ps); }
        else if (k == "AreaLightSource") {
            std::string impl = quoted(); Params ps = parameter_list();
            if (impl == "diffuse") {
                std::string key; Arg a;
                if (!ps.extract_substr("L", key, a) || a.kind != Arg::Numbers) throw unsupported("default / complicated luminance for diffuse light");
                area_l = constant_color(key, a.numbers); has_area_l = true;
            }
        } else if (k == "Texture") {
            std::string name = quoted(), ttype = quoted(), impl = quoted();
            Params ps = parameter_list();
            if (ttype == "color" || ttype == "spectrum") {
                if (impl != "imagemap") throw unsupported("tex impl = " + impl);
                Arg fn;
                if (!ps.extract("string filename", fn) || fn.kind != Arg::String) throw Error(PBRS_ERR_INVALID_ARG, "missing file name for image map texture");
                if (!load_image) throw unsupported("no image decoder installed (Loader::load_image)");
                named_tex[name] = load_image(root + "/" + fn.str);
            }
        } else if (k == "MakeNamedMaterial") {
            std::string name = quoted(); Params ps = parameter_list(); Arg ty;
            if (!ps.extract("string type", ty) || ty.kind != Arg::String) throw Error(PBRS_ERR_INVALID_ARG, "no material type specified");
            named_mtl[name] = material(ty.str, ps);
        } else if (k == "NamedMaterial") { auto it = named_mtl.find(quoted()); cur_mtl = it == named_mtl.end() ? nullptr : it->second; }
        else if (k == "AttributeBegin") {
            ctm.push_back(ctm.back()); cur_mtl = nullptr; has_area_l = false;  // :199-204
            while (starts_world_item(peek())) world_item();
            expect_kw("AttributeEnd"); ctm.pop_back();
        } else if (k == "TransformBegin") {
            ctm.push_back(ctm.back());
            while (starts_world_item(peek())) world_item();
            expect_kw("TransformEnd"); ctm.pop_back();
        } else if (k == "ReverseOrientation") {
        } else throw unsupported("world item " + k + " (object instancing is unimplemented!() upstream, loader.rs:781)");
    }

    void light_source(const std::string &impl, Params &ps) {  // :257-284, :436-481
        std::string key; Arg a;
        if (impl == "infinite") {
            bool has_l = ps.extract_substr("L", key, a);
            Color mult = has_l ? constant_color(key, a.numbers) : Color::white();
            Arg map;
            if (ps.extract("string mapname", map)) {
                if (!load_image) throw unsupported("no image decoder installed (Loader::load_image)");
                env_kind = 2; env_image = load_image(root + "/" + map.str); env_scale = mult;
            } else if (has_l) { env_kind = 0; env_color = mult; }
            else throw Error(PBRS_ERR_INVALID_ARG, "can't process the infinite light");
            return;
        }
        auto pt = [&](const char *name, Point3 dflt) { std::string kk; Arg v; if (!ps.extract_substr(name, kk, v)) return dflt; return Point3{v.numbers.at(0), v.numbers.at(1), v.numbers.at(2)}; };
        if (impl == "distant") {
            Point3 from = pt("from", {0, 0, 0}), to = pt("to", {0, 0, 1});
            Color L = color_arg(ps, "L", Color::white());
            delta.push_back(light::DeltaLight::distant(INFINITY, {to.x - from.x, to.y - from.y, to.z - from.z}, L));  // radius fixed at commit (scene/src/lib.rs:54-58)
        } else if (impl == "point") {
            Point3 from = pt("from", {0, 0, 0});
            delta.push_back(light::DeltaLight::point(from, color_arg(ps, "L", Color::white())));
        } else throw unsupported("light of " + impl);
    }

    Scene scene() {  // parser.rs:21-36 + loader.rs:91-160
        bool has_fov = false, has_pose = false;
        float fov = 60.0f, w = 0, h = 0;
        bool has_w = false, has_h = false;
        std::vector<float> pose;
        InstanceTransform world = InstanceTransform::identity();
        while (peek().kind == Token::Kw && !is_kw("WorldBegin")) {
            if (starts_transform(peek())) {
                Xf x = parse_transform();
                if (x.kind == "LookAt") { pose = x.n; has_pose = true; } else world = world * to_affine(x);
                continue;
            }
            std::string k = t[i++].text;
            if (k != "Camera" && k != "Sampler" && k != "Film" && k != "Filter" && k != "Integrator" && k != "Accelerator")
                throw Error(PBRS_ERR_INVALID_ARG, "incorrect token to start a scene option: " + k);
            std::string impl = quoted();
            Params ps = parameter_list();
            if (k == "Camera") { Arg a; has_fov = true; if (ps.extract("float fov", a)) { if (a.kind != Arg::Number) throw Error(PBRS_ERR_INVALID_ARG, "complicated fov degree"); fov = a.number; } }
            if (k == "Film") { has_w = ps.lookup_f32("integer xresolution", w); has_h = ps.lookup_f32("integer yresolution", h); }
        }
        expect_kw("WorldBegin");
        while (starts_world_item(peek())) world_item();
        expect_kw("WorldEnd");
        if (!has_fov || !has_w || !has_h) throw Error(PBRS_ERR_STATE, "the scene file needs Camera, Film xresolution and yresolution (the reference unwraps None)");
        Camera cam({(uint32_t)w, (uint32_t)h}, Angle::new_deg(fov));
        if (has_pose) cam.look_at({pose[0], pose[1], pose[2]}, {pose[3], pose[4], pose[5]}, {pose[6], pose[7], pose[8]});
        for (Instance &in : instances) in = in.with_transform(world * in.transform);  // :158-160
        Scene sc = Scene(std::move(instances), cam).with_lights(std::move(delta), std::move(area));
        if (env_kind == 2) return std::move(sc).with_env_map(env_image, env_scale);
        return std::move(sc).with_const_env_light(env_color);
    }
};

// scene::loader::build_scene(path) + Scene::from_loader (scene/src/loader.rs:41-58, scene/src/lib.rs:46-63)
inline Scene build_scene_from_string(const std::string &text, const std::string &root_dir = ".", std::function<TextureRef(const std::string &)> load_image = nullptr) {
    Loader l;
    l.root = root_dir;
    l.load_image = std::move(load_image);
    tokenize(text, root_dir, l.t);
    l.t.push_back({Token::End, "$"});
    return l.scene();
}
inline Scene build_scene(const std::string &path, std::function<TextureRef(const std::string &)> load_image = nullptr) {
    return build_scene_from_string(read_file(path), dir_of(path), std::move(load_image));
}

}  // namespace scene_file
}  // namespace pbrs
