// host.cpp — host-side mirror of the reference's config / scene set-up / output layer, in C++ above the
// engine's C ABI (the reference's host is Fortran; no Fortran toolchain exists in this image).
//   parse_params        src/parse/parse.f90:20-72  (+ parse_source/geometry/detectors, App. C of SURVEY.md)
//   setup_simulation    src/setup.f90:14-62 -> src/setupGeometry.f90 (scene contents, App. E)
//   finalise / writers  src/kernelsMod.f90:2321-2416, src/writer.f90
//   default_MCRT        src/kernelsMod.f90:29-83
#include <sys/stat.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/smcrt_host.h"
#include "toml_lite.hpp"

namespace {

using toml_lite::Table;
using toml_lite::Value;

thread_local std::string g_host_err;
int fail(const std::string& m) {
    g_host_err = m;
    return -1;
}

// ------------------------------------------------------------------ small matrix kit (Fortran 4x4 layout)
struct Mat {
    double m[16];  // (i,j) 1-based at m[(j-1)*4+(i-1)]
    double& a(int i, int j) { return m[(j - 1) * 4 + (i - 1)]; }
    double a(int i, int j) const { return m[(j - 1) * 4 + (i - 1)]; }
};
Mat mat_identity() {
    Mat r{};
    for (int i = 1; i <= 4; ++i) r.a(i, i) = 1.0;
    return r;
}
// translate(o): o in row 4 (src/sdfs/sdfHelpers.f90:168-182)
Mat mat_translate(double x, double y, double z) {
    Mat r = mat_identity();
    r.a(4, 1) = x;
    r.a(4, 2) = y;
    r.a(4, 3) = z;
    return r;
}
// rotate_y(angle in degrees) (src/sdfs/sdfHelpers.f90:43-62)
Mat mat_rotate_y(double deg) {
    const double a = deg * 3.14159265358979323846 / 180.0, c = std::cos(a), s = std::sin(a);
    Mat r = mat_identity();
    r.a(1, 1) = c;  r.a(3, 1) = s;
    r.a(1, 3) = -s; r.a(3, 3) = c;
    return r;
}
// invert (src/mat_class.f90:154-214 is a closed form); Gauss-Jordan with partial pivoting here
Mat mat_invert(const Mat& in) {
    double w[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            w[i][j] = in.a(i + 1, j + 1);
            w[i][j + 4] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
            if (std::fabs(w[r][c]) > std::fabs(w[piv][c])) piv = r;
        if (piv != c)
            for (int j = 0; j < 8; ++j) std::swap(w[c][j], w[piv][j]);
        double d = w[c][c];
        for (int j = 0; j < 8; ++j) w[c][j] /= d;
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                double f = w[r][c];
                if (f != 0.0)
                    for (int j = 0; j < 8; ++j) w[r][j] -= f * w[c][j];
            }
    }
    Mat out{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out.a(i + 1, j + 1) = w[i][j + 4];
    return out;
}

// ------------------------------------------------------------------ scene tree -> flat node table
struct SdfTree {
    int kind = 0;
    Mat xf = mat_identity();
    double p[SMCRT_NODE_PARAMS] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<SdfTree> kids;
};
struct TopSdf {
    SdfTree tree;
    double mus, mua, hgg, n;
};
struct FlatScene {
    std::vector<int32_t> kind, first_child, n_child, top_node;
    std::vector<double> xform, params, mus, mua, hgg, n;
    int add_slot() {
        kind.push_back(0);
        first_child.push_back(0);
        n_child.push_back(0);
        xform.resize(xform.size() + 16, 0.0);
        params.resize(params.size() + SMCRT_NODE_PARAMS, 0.0);
        return (int)kind.size() - 1;
    }
    void fill(int idx, const SdfTree& t) {
        kind[idx] = t.kind;
        std::memcpy(&xform[16 * (size_t)idx], t.xf.m, sizeof(double) * 16);
        std::memcpy(&params[SMCRT_NODE_PARAMS * (size_t)idx], t.p, sizeof(double) * SMCRT_NODE_PARAMS);
        n_child[idx] = (int32_t)t.kids.size();
        if (!t.kids.empty()) {
            int first = -1;
            for (size_t k = 0; k < t.kids.size(); ++k) {
                int s = add_slot();
                if (k == 0) first = s;
            }
            first_child[idx] = first;
            for (size_t k = 0; k < t.kids.size(); ++k) fill(first + (int)k, t.kids[k]);
        }
    }
    void add_top(const TopSdf& t) {
        int idx = add_slot();
        top_node.push_back(idx);
        fill(idx, t.tree);
        mus.push_back(t.mus);
        mua.push_back(t.mua);
        hgg.push_back(t.hgg);
        n.push_back(t.n);
    }
};

SdfTree prim_sphere(double r, const Mat* xf = nullptr) {
    SdfTree t;
    t.kind = SMCRT_SPHERE;
    t.p[0] = r;
    if (xf) t.xf = *xf;
    return t;
}
// box(lengths): stores HALF lengths (src/sdfs/sdfs.f90:455)
SdfTree prim_box(double lx, double ly, double lz, const Mat* xf = nullptr) {
    SdfTree t;
    t.kind = SMCRT_BOX;
    t.p[0] = 0.5 * lx; t.p[1] = 0.5 * ly; t.p[2] = 0.5 * lz;
    if (xf) t.xf = *xf;
    return t;
}
SdfTree prim_cylinder(const double a[3], const double b[3], double r, const Mat* xf = nullptr) {
    SdfTree t;
    t.kind = SMCRT_CYLINDER;
    for (int i = 0; i < 3; ++i) { t.p[i] = a[i]; t.p[3 + i] = b[i]; }
    t.p[6] = r;
    if (xf) t.xf = *xf;
    return t;
}
SdfTree prim_capsule(const double a[3], const double b[3], double r) {
    SdfTree t;
    t.kind = SMCRT_CAPSULE;
    for (int i = 0; i < 3; ++i) { t.p[i] = a[i]; t.p[3 + i] = b[i]; }
    t.p[6] = r;
    return t;
}
SdfTree prim_torus(double R, double r, const Mat* xf = nullptr) {
    SdfTree t;
    t.kind = SMCRT_TORUS;
    t.p[0] = R; t.p[1] = r;
    if (xf) t.xf = *xf;
    return t;
}
SdfTree prim_egg(double r1, double r2, double h) {
    SdfTree t;
    t.kind = SMCRT_EGG;
    t.p[0] = r1; t.p[1] = r2; t.p[2] = h;
    return t;
}
SdfTree mod_revolution(const SdfTree& prim, double o, double cx, double cy, double cz) {
    SdfTree t;
    t.kind = SMCRT_MOD_REVOLUTION;
    t.p[0] = o; t.p[1] = cx; t.p[2] = cy; t.p[3] = cz;
    t.kids.push_back(prim);
    return t;
}
SdfTree model_of(int op, const std::vector<SdfTree>& kids, double k) {
    SdfTree t;
    t.kind = op;
    t.p[0] = k;
    t.kids = kids;
    return t;
}

// ------------------------------------------------------------------ config = `state` + `dict` + dects
struct DetCfg {
    int kind = 0;
    std::string id;
    double p[SMCRT_DET_PARAMS] = {0};
    int nbins = 100;
    int layer = 1;
    bool track = false;  // trackHistory
};
struct Config {
    // [source]  (parse_source.f90:58-255)
    std::string source = "point";
    int64_t nphotons = 1000000;
    double src[SMCRT_SOURCE_PARAMS] = {0};
    int src_kind = SMCRT_SRC_POINT, src_subtype = 0;
    std::string annulus_type = "gaussian", focus_type = "gaussian";
    double wavelength = 500.0;
    // [grid]  (parse.f90:92-110)
    int nxg = 200, nyg = 200, nzg = 200;
    double xmax = 1.0, ymax = 1.0, zmax = 1.0;
    std::string units = "cm";
    // [geometry]  (parse_geometry.f90:45-282)
    std::string geom = "sphere";
    int numOptProp = 1, num_spheres = 10;
    std::vector<double> mua, mus, mur, hgg, nref;
    double position[3] = {0, 0, 0}, boundingBox[3] = {2, 2, 2}, BoxDimensions[3] = {1, 1, 1};
    double sphereRadius = 1.0, tau = 10.0;
    double musb = 0.0, muab = 0.01, musc = 0.0, muac = 0.01, hgga = 0.7;
    double BottomSphereRadius = 3.0, TopSphereRadius = 0, SphereSep = 0, ShellThickness = 0.05, YolkRadius = 1.5;
    // [[detectors]]
    std::vector<DetCfg> dets;  // already in dects(:) order
    // [output] (parse.f90:127-155)
    std::string outfile = "fluence.nrrd", outfile_absorb = "absorb.nrrd", rendergeomfile = "geom_render.nrrd",
                rendersourcefile = "source_render.nrrd";
    bool render_geom = false, render_source = false, overwrite = false;
    // [simulation] (parse.f90:170-184)
    int64_t iseed = 123456789;
    bool tev = false, absorb = false, loadckpt = false;
    std::string history_file = "photPos.obj";  // [[detectors]] historyFileName
    std::string ckptfile = "check.ckpt", ckpt_deck;  // ckpt_deck: the input deck's name as written into checkpoints
    int64_t ckptfreq = 1000000;
    // derived
    std::string res_dir;
    FlatScene scene;
    std::vector<std::pair<std::string, std::string>> dict;  // metadata in insertion order
    std::string meta_text;
    void dict_set(const std::string& k, const std::string& v) {
        for (auto& kv : dict)
            if (kv.first == k) { kv.second = v; return; }
        dict.emplace_back(k, v);
    }
    void dict_set(const std::string& k, double v) {
        char b[64];
        std::snprintf(b, sizeof b, "%.17g", v);
        std::string s = b;
        if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
        dict_set(k, s);
    }
    void dict_seti(const std::string& k, long long v) { dict_set(k, std::to_string(v)); }
    void dict_sets(const std::string& k, const std::string& v) { dict_set(k, "\"" + v + "\""); }
};

struct CfgError {
    std::string msg;
};
[[noreturn]] void cfg_fail(const std::string& m) { throw CfgError{m}; }

const Value* find(const Table& t, const std::string& k) {
    auto it = t.find(k);
    return it == t.end() ? nullptr : &it->second;
}
double get_num(const Table& t, const std::string& k, double def) {
    const Value* v = find(t, k);
    if (!v) return def;
    if (!v->is_number()) cfg_fail("key '" + k + "' must be a number");
    return v->as_double();
}
long long get_int(const Table& t, const std::string& k, long long def) {
    const Value* v = find(t, k);
    if (!v) return def;
    if (v->kind == Value::INT) return v->i;
    if (v->kind == Value::FLOAT && v->f == std::floor(v->f)) return (long long)v->f;
    cfg_fail("key '" + k + "' must be an integer");
}
bool get_bool(const Table& t, const std::string& k, bool def) {
    const Value* v = find(t, k);
    if (!v) return def;
    if (v->kind != Value::BOOL) cfg_fail("key '" + k + "' must be a boolean");
    return v->b;
}
std::string get_str(const Table& t, const std::string& k, const std::string& def) {
    const Value* v = find(t, k);
    if (!v) return def;
    if (v->kind != Value::STRING) cfg_fail("key '" + k + "' must be a string");
    return v->s;
}
// get_vector (src/parse/parse_helpers.f90): a 3-array of numbers; returns false if absent / not an array
bool get_vec3(const Table& t, const std::string& k, double out[3]) {
    const Value* v = find(t, k);
    if (!v || v->kind != Value::ARRAY) return false;
    if (v->arr.size() != 3) cfg_fail("'" + k + "': expected vector of size 3");
    for (int i = 0; i < 3; ++i) {
        if (!v->arr[i].is_number()) cfg_fail("'" + k + "': expected numbers");
        out[i] = v->arr[i].as_double();
    }
    return true;
}
void normalise3(double v[3]) {
    double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int i = 0; i < 3; ++i) v[i] /= l;
}
std::string i4(int i) {  // Fortran write(string,'(I4)') i
    char b[16];
    std::snprintf(b, sizeof b, "%4d", i);
    return b;
}

// ---- [source]  (src/parse/parse_source.f90:17-264)
void parse_source(const Table& root, Config& c) {
    const Value* sv = find(root, "source");
    if (!sv || sv->kind != Value::TABLE) cfg_fail("Simulation needs Source table");
    const Table& t = *sv->tbl;
    const char* axis[3] = {"x", "y", "z"};
    c.source = get_str(t, "name", "point");
    c.nphotons = get_int(t, "nphotons", 1000000);
    static const std::pair<const char*, int> kinds[] = {{"point", SMCRT_SRC_POINT},     {"pencil", SMCRT_SRC_PENCIL},
                                                         {"uniform", SMCRT_SRC_UNIFORM}, {"circular", SMCRT_SRC_CIRCULAR},
                                                         {"focus", SMCRT_SRC_FOCUS},     {"annulus", SMCRT_SRC_ANNULUS},
                                                         {"dslit", SMCRT_SRC_DSLIT},     {"aperture", SMCRT_SRC_APERTURE}};
    c.src_kind = 0;
    for (auto& k : kinds)
        if (c.source == k.first) c.src_kind = k.second;
    if (c.src_kind == 0) {
        if (c.source == "slm")
            cfg_fail("source 'slm' (image-driven emitter, src/photon.f90:159-212) is outside the hot-path scope of this engine");
        cfg_fail("No such source!");  // init_source, src/photon.f90:155
    }
    double pos[3] = {0, 0, 0}, dir[3] = {0, 0, 0}, rot[3] = {0, 0, 0};
    if (c.source != "uniform") {
        if (!get_vec3(t, "position", pos)) cfg_fail("source needs a 'position' vector");
    }
    if (c.source == "focus" || c.source == "annulus" || c.source == "dslit" || c.source == "aperture") {
        if (!get_vec3(t, "rotation", rot)) cfg_fail("Source requires rotation variable");
        double l = std::sqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]);
        if (l < 1e-8) cfg_fail("Need to specify rotation that has length greater than 0.0");
        normalise3(rot);
        for (int i = 0; i < 3; ++i) c.dict_set(std::string("rotation%") + axis[i], rot[i]);
    }
    // direction: vector, or a cardinal string.  (The reference returns early from parse_source when the
    // direction is given as a VECTOR, parse_source.f90:153-169, leaving the source half-initialised; that is a
    // parser bug, not path behaviour — vectors are honoured here.)
    bool have_dir = get_vec3(t, "direction", dir);
    if (!have_dir) {
        const Value* dv = find(t, "direction");
        if (dv && dv->kind == Value::STRING) {
            const std::string& d = dv->s;
            if (d == "x") dir[0] = 1;
            else if (d == "-x") dir[0] = -1;
            else if (d == "y") dir[1] = 1;
            else if (d == "-y") dir[1] = -1;
            else if (d == "z") dir[2] = 1;
            else if (d == "-z") dir[2] = -1;
            else cfg_fail("Direction needs a cardinal direction i.e x, y, or z");
            have_dir = true;
        } else if (c.source != "point" && c.source != "annulus" && c.source != "focus") {
            cfg_fail("Need to specify direction for source type!");
        }
    }
    // corners default (parse_source.f90:52-56) then point1..3
    double corners[3][3] = {{-1, -1, 1}, {2, 0, 0}, {0, 2, 0}};
    const char* pkeys[3] = {"point1", "point2", "point3"};
    for (int k = 0; k < 3; ++k) {
        double v[3];
        const Value* pv = find(t, pkeys[k]);
        if (pv && pv->kind == Value::ARRAY) {
            if (pv->arr.size() < 3) cfg_fail("Need a matrix row for points");
            for (int i = 0; i < 3; ++i) v[i] = pv->arr[i].as_double();
            for (int i = 0; i < 3; ++i) {
                corners[k][i] = v[i];
                c.dict_set("pos" + std::to_string(k + 1) + "%" + axis[i], v[i]);
            }
        } else if (c.source == "uniform")
            cfg_fail(std::string("Uniform source requires ") + pkeys[k] + " variable");
    }
    double radius = get_num(t, "radius", 0.5);          c.dict_set("radius", radius);
    double focal = get_num(t, "focalLength", 1.0);      c.dict_set("focalLength", focal);
    double rhi = get_num(t, "rhi", 0.6);                c.dict_set("rhi", rhi);
    double rlo = get_num(t, "rlo", 0.5);                c.dict_set("rlo", rlo);
    double sigma = get_num(t, "sigma", 0.04);           c.dict_set("sigma", sigma);
    c.annulus_type = get_str(t, "annulus_type", "gaussian"); c.dict_sets("annulus_type", c.annulus_type);
    c.focus_type = get_str(t, "focus_type", "gaussian");     c.dict_sets("focus_type", c.focus_type);
    double beam = get_num(t, "beam_size", 0.5);         c.dict_set("beam_size", beam);
    // spectrum: only "constant" is on the hot path (parse_spectrum.f90:52-117)
    std::string st = get_str(t, "spectrum_type", "constant");
    if (st != "constant") cfg_fail("spectrum_type '" + st + "' (tabulated spectra) is outside the hot-path scope");
    c.wavelength = get_num(t, "wavelength", 500.0);
    c.dict_sets("spectrum_type", st);
    c.dict_set("wavelength", c.wavelength);

    double* s = c.src;
    for (int i = 0; i < 3; ++i) {
        s[SMCRT_SP_POS + i] = pos[i];
        s[SMCRT_SP_DIR + i] = dir[i];
        s[SMCRT_SP_P1 + i] = corners[0][i];
        s[SMCRT_SP_P2 + i] = corners[1][i];
        s[SMCRT_SP_P3 + i] = corners[2][i];
        s[SMCRT_SP_ROT + i] = rot[i];
    }
    s[SMCRT_SP_RADIUS] = (c.src_kind == SMCRT_SRC_DSLIT || c.src_kind == SMCRT_SRC_APERTURE) ? c.wavelength : radius;  // (see smcrt.h)
    s[SMCRT_SP_FOCAL] = focal; s[SMCRT_SP_BEAM] = beam;
    s[SMCRT_SP_RLO] = rlo; s[SMCRT_SP_RHI] = rhi; s[SMCRT_SP_SIGMA] = sigma;
    c.src_subtype = 0;
    if (c.src_kind == SMCRT_SRC_FOCUS) {
        if (c.focus_type == "square") c.src_subtype = SMCRT_FOCUS_SQUARE;
        else if (c.focus_type == "circle") c.src_subtype = SMCRT_FOCUS_CIRCLE;
        else if (c.focus_type == "gaussian") c.src_subtype = SMCRT_FOCUS_GAUSSIAN;
        else cfg_fail("No such beam type!");
    } else if (c.src_kind == SMCRT_SRC_ANNULUS) {
        if (c.annulus_type == "tophat") c.src_subtype = SMCRT_ANNULUS_TOPHAT;
        else if (c.annulus_type == "besselAnnulus") c.src_subtype = SMCRT_ANNULUS_BESSEL;
        else if (c.annulus_type == "gaussian") c.src_subtype = SMCRT_ANNULUS_GAUSSIAN;
        else cfg_fail("No such beam type!");
    }
}

// ---- [grid]  (src/parse/parse.f90:75-123)
void parse_grid(const Table& root, Config& c) {
    const Value* gv = find(root, "grid");
    if (!gv || gv->kind != Value::TABLE) cfg_fail("Need grid table in input param file");
    const Table& t = *gv->tbl;
    c.nxg = (int)get_int(t, "nxg", 200);
    c.nyg = (int)get_int(t, "nyg", 200);
    c.nzg = (int)get_int(t, "nzg", 200);
    c.xmax = get_num(t, "xmax", 1.0);
    c.ymax = get_num(t, "ymax", 1.0);
    c.zmax = get_num(t, "zmax", 1.0);
    c.units = get_str(t, "units", "cm");
    c.dict_sets("units", c.units);
}

// ---- [geometry]  (src/parse/parse_geometry.f90:17-292)
void parse_geometry(const Table& root, Config& c) {
    const Value* gv = find(root, "geometry");
    if (!gv || gv->kind != Value::TABLE) cfg_fail("Need geometry table in input param file");
    const Table& t = *gv->tbl;
    c.geom = get_str(t, "geom_name", "sphere");
    c.tau = get_num(t, "tau", 10.0);                      c.dict_set("tau", c.tau);
    c.num_spheres = (int)get_int(t, "num_spheres", 10);   c.dict_seti("num_spheres", c.num_spheres);
    c.musb = get_num(t, "musb", 0.0);   c.dict_set("musb", c.musb);
    c.muab = get_num(t, "muab", 0.01);  c.dict_set("muab", c.muab);
    c.musc = get_num(t, "musc", 0.0);   c.dict_set("musc", c.musc);
    c.muac = get_num(t, "muac", 0.01);  c.dict_set("muac", c.muac);
    c.hgga = get_num(t, "hgga", 0.7);   c.dict_set("hgga", c.hgga);
    c.numOptProp = (int)get_int(t, "numOptProp", 1);      c.dict_seti("numOptProp", c.numOptProp);
    if (c.numOptProp < 1) cfg_fail("Need to set an integer value of at least one or greater for numOptProp");
    if (c.geom == "sphere" && c.numOptProp != 1) cfg_fail("For geometry of sphere must set numOptProp to one");
    if (c.geom == "box" && c.numOptProp != 1) cfg_fail("For geometry of box must set numOptProp to one");
    if (c.geom == "egg" && c.numOptProp != 3) cfg_fail("For geometry of egg must set numOptProp to three");
    struct Arr { const char* key; std::vector<double>* dst; double def; };
    Arr arrs[5] = {{"mua", &c.mua, 0.0}, {"mus", &c.mus, 1.0}, {"mur", &c.mur, 0.0}, {"hgg", &c.hgg, 0.0}, {"n", &c.nref, 1.0}};
    for (auto& a : arrs) {
        a.dst->assign(c.numOptProp, a.def);
        const Value* v = find(t, a.key);
        if (v && v->kind == Value::ARRAY) {
            if ((int)v->arr.size() != c.numOptProp) cfg_fail(std::string("length of ") + a.key + " must be equal to numOptProp");
            for (int i = 0; i < c.numOptProp; ++i) (*a.dst)[i] = v->arr[i].as_double();
        }
        for (int i = 0; i < c.numOptProp; ++i) c.dict_set(std::string(a.key) + "%" + i4(i + 1), (*a.dst)[i]);
    }
    get_vec3(t, "position", c.position);
    get_vec3(t, "boundingBox", c.boundingBox);
    for (int i = 0; i < 3; ++i) c.dict_set("position%" + i4(i + 1), c.position[i]);
    for (int i = 0; i < 3; ++i) c.dict_set("boundinglength%" + i4(i + 1), c.boundingBox[i]);
    c.sphereRadius = get_num(t, "sphereRadius", 1.0);  c.dict_set("sphereRadius", c.sphereRadius);
    get_vec3(t, "BoxDimensions", c.BoxDimensions);
    for (int i = 0; i < 3; ++i) c.dict_set("BoxDimensions%" + i4(i + 1), c.BoxDimensions[i]);
    const double d = 3.0 * std::sqrt(2.0 - std::sqrt(2.0));
    c.BottomSphereRadius = get_num(t, "BottomSphereRadius", 3.0);
    c.TopSphereRadius = get_num(t, "TopSphereRadius", d);
    c.SphereSep = get_num(t, "SphereSep", d);
    c.ShellThickness = get_num(t, "ShellThickness", 0.05);
    c.YolkRadius = get_num(t, "YolkRadius", 1.5);
}

// ---- [[detectors]]  (src/parse/parse_detectors.f90:17-349)
void parse_detectors(const Table& root, Config& c) {
    const Value* dv = find(root, "detectors");
    if (!dv) return;
    if (dv->kind != Value::TABLE_ARRAY) cfg_fail("'detectors' must be an array of tables");
    std::vector<DetCfg> circ, ann, fib, cam;
    for (auto& tp : dv->tarr) {
        const Table& t = *tp;
        const Value* tv = find(t, "type");
        if (!find(t, "ID")) cfg_fail("Need to specify a detector ID");
        std::string type = tv && tv->kind == Value::STRING ? tv->s : "";
        DetCfg d;
        d.id = get_str(t, "ID", "none");
        d.layer = (int)get_int(t, "layer", 1);
        // (the parallel reference build refuses trackHistory, parse_detectors.f90:178-181; the engine records the hits and replays
        // those packets, include/smcrt.h)
        d.track = get_bool(t, "trackHistory", false);
        if (const Value* hv = find(t, "historyFileName"); hv && hv->kind == Value::STRING) c.history_file = hv->s;
        double pos[3], dir[3] = {0, 0, -1};
        if (type == "circle") {
            if (!get_vec3(t, "position", pos)) cfg_fail("detector needs a position");
            get_vec3(t, "direction", dir);
            normalise3(dir);
            d.kind = SMCRT_DET_CIRCLE;
            for (int i = 0; i < 3; ++i) { d.p[i] = pos[i]; d.p[3 + i] = dir[i]; }
            d.p[6] = get_num(t, "radius", 1.0);
            d.nbins = (int)get_int(t, "nbins", 100);
            circ.push_back(d);
        } else if (type == "annulus") {
            if (!get_vec3(t, "position", pos)) cfg_fail("detector needs a position");
            get_vec3(t, "direction", dir);  // NOT normalised in the reference (parse_detectors.f90:318)
            d.kind = SMCRT_DET_ANNULUS;
            for (int i = 0; i < 3; ++i) { d.p[i] = pos[i]; d.p[3 + i] = dir[i]; }
            d.p[6] = get_num(t, "radius1", 0.1);
            d.p[7] = get_num(t, "radius2", 0.2);
            if (d.p[7] <= d.p[6]) cfg_fail("Radii are invalid");
            d.nbins = (int)get_int(t, "nbins", 100);
            ann.push_back(d);
        } else if (type == "fibre") {
            if (!get_vec3(t, "position", pos)) cfg_fail("detector needs a position");
            get_vec3(t, "direction", dir);
            normalise3(dir);
            d.kind = SMCRT_DET_FIBRE;
            for (int i = 0; i < 3; ++i) { d.p[i] = pos[i]; d.p[3 + i] = dir[i]; }
            double f1 = get_num(t, "focalLength1", 1.0), f2 = get_num(t, "focalLength2", 1.0);
            double a1 = get_num(t, "f1Aperture", 1.0), a2 = get_num(t, "f2Aperture", 1.0);
            d.p[6] = f1; d.p[7] = f2; d.p[8] = a1; d.p[9] = a2;
            d.p[10] = get_num(t, "frontOffset", 0.0);
            d.p[11] = get_num(t, "backOffset", f2);
            d.p[12] = get_num(t, "frontToPinSep", f1);
            d.p[13] = get_num(t, "pinToBackSep", f2);
            d.p[14] = get_num(t, "pinAperture", std::max(a1, a2));
            d.p[15] = get_num(t, "acceptanceAngle", 90.0);  // the shipped tomls write `acceptAngle`, which is ignored
            d.p[16] = get_num(t, "coreDiameter", 0.01);
            d.nbins = (int)get_int(t, "nbins", 1);
            fib.push_back(d);
        } else if (type == "camera") {
            double p1[3] = {-1, -1, -1}, p2[3] = {2, 0, 0}, p3[3] = {0, 2, 0};
            get_vec3(t, "p1", p1);
            get_vec3(t, "p2", p2);
            get_vec3(t, "p3", p3);
            d.kind = SMCRT_DET_CAMERA;
            for (int i = 0; i < 3; ++i) { d.p[i] = p1[i]; d.p[3 + i] = p2[i]; d.p[6 + i] = p3[i]; }
            d.p[9] = get_num(t, "maxval", 100.0);
            d.nbins = (int)get_int(t, "nbins", 100);
            cam.push_back(d);
        } else
            cfg_fail("Invalid detector type. Valid types are [circle, annulus, camera]");
    }
    // dects(:) order: circles, annuli, fibres, cameras (parse_detectors.f90:119-137)
    for (auto* v : {&circ, &ann, &fib, &cam})
        for (auto& d : *v) c.dets.push_back(d);
}

void parse_output_sim(const Table& root, Config& c) {
    if (const Value* ov = find(root, "output"); ov && ov->kind == Value::TABLE) {
        const Table& t = *ov->tbl;
        c.outfile = get_str(t, "fluence", "fluence.nrrd");
        c.outfile_absorb = get_str(t, "absorb", "absorb.nrrd");
        c.rendergeomfile = get_str(t, "render_geometry_name", "geom_render.nrrd");
        c.render_geom = get_bool(t, "render_geometry", false);
        c.rendersourcefile = get_str(t, "render_source_name", "source_render.nrrd");
        c.render_source = get_bool(t, "render_source", false);
        c.overwrite = get_bool(t, "overwrite", false);
    }
    if (const Value* sv = find(root, "simulation"); sv && sv->kind == Value::TABLE) {
        const Table& t = *sv->tbl;
        c.iseed = get_int(t, "iseed", 123456789);
        c.tev = get_bool(t, "tev", false);
        c.absorb = get_bool(t, "absorb", false);
        c.loadckpt = get_bool(t, "load_checkpoint", false);
        c.ckptfile = get_str(t, "checkpoint_file", "check.ckpt");
        c.ckptfreq = get_int(t, "checkpoint_every_n", 1000000);
    }
}

// ------------------------------------------------------------------ scene builders (src/setupGeometry.f90)
struct SplitMix {  // seeded stand-in for the reference's UNSEEDED ranu() in setup_sphere_scene
    uint64_t s;
    double uni() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
    double ranu(double a, double b) { return a + uni() * (b - a); }
};

bool read_table(const std::string& path, int ncol, std::vector<double>& out) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::vector<double> row;
        double v;
        while ((int)row.size() < ncol && ss >> v) row.push_back(v);
        if ((int)row.size() < ncol) break;  // iostat /= 0 -> stop, like the reference's read loop
        out.insert(out.end(), row.begin(), row.end());
    }
    return true;
}

void build_scene(Config& c) {
    FlatScene& S = c.scene;
    S = FlatScene();
    const std::string& g = c.geom;
    auto mus1 = [&](int i) { return c.mus.at(i); };
    if (g == "sphere") {  // setup_sphere :10-71
        Mat t = mat_invert(mat_translate(c.position[0], c.position[1], c.position[2]));
        S.add_top({prim_sphere(c.sphereRadius, &t), mus1(0), c.mua[0], c.hgg[0], c.nref[0]});
        S.add_top({prim_box(c.boundingBox[0], c.boundingBox[1], c.boundingBox[2]), 0.0, 0.0, 0.0, 1.0});
    } else if (g == "box" || g == "test_box") {  // setup_box :73-147
        Mat t = mat_invert(mat_translate(c.position[0], c.position[1], c.position[2]));
        S.add_top({prim_box(c.BoxDimensions[0], c.BoxDimensions[1], c.BoxDimensions[2], &t), mus1(0), c.mua[0], c.hgg[0], c.nref[0]});
        S.add_top({prim_box(c.boundingBox[0], c.boundingBox[1], c.boundingBox[2]), 0.0, 0.0, 0.0, 1.0});
    } else if (g == "egg") {  // setup_egg :149-248  (array order: yolk, albumen, shell, bbox)
        if (c.numOptProp < 3) cfg_fail("For geometry of egg must set numOptProp to three");
        Mat t = mat_invert(mat_translate(c.position[0], c.position[1], c.position[2]));
        const double k = 1.0 - c.ShellThickness;
        SdfTree shell = mod_revolution(prim_egg(c.BottomSphereRadius, c.TopSphereRadius, c.SphereSep), 0.0, c.position[0],
                                       c.position[1], c.position[2]);
        SdfTree albumen = mod_revolution(prim_egg(c.BottomSphereRadius * k, c.TopSphereRadius * k, c.SphereSep * k), 0.0,
                                         c.position[0], c.position[1], c.position[2]);
        S.add_top({prim_sphere(c.YolkRadius, &t), c.mus[2], c.mua[2], c.hgg[2], c.nref[2]});
        S.add_top({albumen, c.mus[1], c.mua[1], c.hgg[1], c.nref[1]});
        S.add_top({shell, c.mus[0], c.mua[0], c.hgg[0], c.nref[0]});
        S.add_top({prim_box(c.boundingBox[0], c.boundingBox[1], c.boundingBox[2]), 0.0, 0.0, 0.0, 1.0});
    } else if (g == "sphere_scene") {  // setup_sphere_scene :250-294
        // The reference draws radii/centres from ranu() BEFORE any init_rng (kernelsMod.f90:2302 is commented
        // out), i.e. from the compiler's default state: not reproducible.  Same formulas and draw order here,
        // from a fixed SplitMix64 stream (seed below); tests/golden/sphere_scene_40.json pins the table.
        SplitMix rng{0x5343454E45343021ull};
        for (int i = 0; i < c.num_spheres; ++i) {
            double radius = rng.ranu(0.001, 0.25);
            double x = rng.ranu(-1.0 + radius, 1.0 - radius);
            double y = rng.ranu(-1.0 + radius, 1.0 - radius);
            double z = rng.ranu(-1.0 + radius, 1.0 - radius);
            Mat t = mat_invert(mat_translate(x, y, z));
            S.add_top({prim_sphere(radius, &t), 0.0, 0.0, 0.9, 1.37});
        }
        S.add_top({prim_box(2, 2, 2), 1e-17, 1e-17, 0.0, 1.0});
    } else if (g == "aptran") {  // setup_tran_and_jacques :335-363
        Mat t = mat_invert(mat_translate(0, 0, 0));
        S.add_top({prim_sphere(0.5, &t), 0.0, 1e-17, 0.0, 1.33});
        S.add_top({prim_box(2, 2, 2), 0.0, 1e-17, 0.0, 1.0});
        S.add_top({prim_box(2.01, 2.01, 2.01), 0.0, 10000000.0, 0.0, 1.0});
    } else if (g == "exp") {  // setup_exp :365-407
        double a[3] = {-8, 0, 0}, b[3] = {8, 0, 0};
        S.add_top({prim_cylinder(a, b, 1.55), c.musc, c.muac, c.hgga, 1.3});
        S.add_top({prim_cylinder(a, b, 1.75), c.musb, c.muab, c.hgga, 1.5});
        S.add_top({prim_box(20, 20, 20), 0.0, 0.0, 0.0, 1.0});
    } else if (g == "scat_test") {  // setup_scat_test :409-435
        S.add_top({prim_sphere(1.0), c.tau, 0.0, 0.0, 1.0});
        S.add_top({prim_box(2, 2, 2), 0.0, 0.0, 0.0, 1.0});
    } else if (g == "scat_test2") {  // setup_scat_test2 :437-464
        S.add_top({prim_box(200, 200, 200), c.tau, 1e-17, c.hgg[0], 1.0});
    } else if (g == "omg") {  // setup_omg_sdf :466-549
        std::vector<SdfTree> k;
        Mat t = mat_invert(mat_translate(0, 0, -0.7));
        k.push_back(prim_torus(0.2, 0.05, &t));
        Mat ry = mat_invert(mat_rotate_y(90.0));
        const double seg[9][6] = {{-.25, 0, -.25, -.25, 0, .25}, {-.25, 0, -.25, .25, 0, .0},   {.25, 0, .0, -.25, 0, .25},
                                  {-.25, 0, .25, .25, 0, .25},   {-.25, 0, .5, .25, 0, .5},     {-.25, 0, .5, -.25, 0, .75},
                                  {.25, 0, .5, .25, 0, .75},     {.25, 0, .75, 0, 0, .75},      {0, 0, .625, 0, 0, .75}};
        for (int i = 0; i < 9; ++i) k.push_back(prim_cylinder(&seg[i][0], &seg[i][3], 0.05, i == 0 ? &ry : nullptr));
        S.add_top({model_of(SMCRT_MODEL_SMOOTHUNION, k, 0.09), 10.0, 0.16, 0.0, 2.65});
        S.add_top({prim_box(2, 2, 2), 0.0, 0.0, 0.0, 1.0});
    } else if (g == "vessels") {  // get_vessels :552-652
        std::vector<double> edges, nodes, radii;
        const std::string d = c.res_dir.empty() ? std::string("res") : c.res_dir;
        if (!read_table(d + "/edges.dat", 2, edges) || !read_table(d + "/nodes.dat", 3, nodes) ||
            !read_table(d + "/radii.dat", 1, radii))
            cfg_fail("vessels geometry needs edges.dat, nodes.dat, radii.dat in '" + d +
                     "' (not shipped by the reference: *.dat is git-ignored; tools/make_vessels.py writes a synthetic tree)");
        const size_t ne = edges.size() / 2, nn = nodes.size() / 3;
        if (ne == 0 || nn == 0 || radii.size() < nn) cfg_fail("vessels data files are empty or inconsistent");
        const double res = 0.001;
        double mx[3] = {0, 0, 0};
        for (size_t i = 0; i < nn; ++i)
            for (int a = 0; a < 3; ++a) mx[a] = std::max(mx[a], std::fabs(nodes[3 * i + a]));
        // NB the reference reads only edge_cnt node rows (:606-609); all node rows are used here.
        for (size_t i = 0; i < nn; ++i)
            for (int a = 0; a < 3; ++a) nodes[3 * i + a] = ((nodes[3 * i + a] / mx[a]) - 0.5) * mx[a] * res;
        for (size_t e = 0; e < ne; ++e) {
            long i1 = (long)edges[2 * e] - 1, i2 = (long)edges[2 * e + 1] - 1;
            if (i1 < 0 || i2 < 0 || (size_t)i1 >= nn || (size_t)i2 >= nn) cfg_fail("vessels edges.dat index out of range");
            S.add_top({prim_capsule(&nodes[3 * i1], &nodes[3 * i2], radii[i1] * res), 94.0, 231.0, 0.9, 1.37});
        }
        S.add_top({prim_box(0.32, 0.18, 0.26), 357.0, 0.458, 0.9, 1.37});
    } else if (g == "jacques" || g == "skin" || g == "lens") {
        // BUILDER-DEFINED geometries.  res/jacques.toml, res/skin.toml and res/lens.toml name geom_names that do not exist
        // in this fork's dispatcher (src/setup.f90:33-60 -> error stop "no such routine"), so there is no reference
        // scene to mirror.  They are defined here (documented in DESIGN.md §7) so that BASELINE configs 2-4 can run; parity
        // for them is engine-vs-oracle only.
        if (g == "jacques") {
            // Jacques' classic semi-infinite-like tissue block: 2^3 cube, mua=1, mus=100, g=0.9, n=1.38, in air.
            S.add_top({prim_box(2, 2, 2), 100.0, 1.0, 0.9, 1.38});
            S.add_top({prim_box(2.02, 2.02, 2.02), 0.0, 0.0, 0.0, 1.0});
        } else if (g == "skin") {
            // five-layer skin model inside the +-0.05 cm cube of res/skin.toml, z from the top face downwards (cm):
            // stratum corneum 0.002, living epidermis 0.008, papillary dermis 0.02, reticular dermis 0.05, hypodermis 0.02;
            // per layer (mus, mua, g, n) at ~630 nm (order-of-magnitude literature values; builder-defined)
            struct L { double t, mus, mua, g, n; };
            const L layers[5] = {{0.002, 1000.0, 0.10, 0.86, 1.50}, {0.008, 450.0, 1.50, 0.80, 1.34}, {0.020, 300.0, 0.70, 0.90, 1.40},
                                 {0.050, 200.0, 0.50, 0.95, 1.39}, {0.020, 150.0, 0.20, 0.75, 1.44}};
            double top = 0.05;
            for (const L& l : layers) {
                Mat t = mat_invert(mat_translate(0.0, 0.0, top - 0.5 * l.t));
                S.add_top({prim_box(0.1, 0.1, l.t, &t), l.mus, l.mua, l.g, l.n});
                top -= l.t;
            }
            S.add_top({prim_box(0.102, 0.102, 0.102), 0.0, 0.0, 0.0, 1.0});  // air shell around the stack
        } else {
            // bi-convex lens: intersection of two spheres of radius 1.0 centred at z = +-0.8 (thickness 0.4, aperture radius 0.6),
            // glass n = 1.5, non-scattering, in the 2^3 air box of res/lens.toml
            Mat ta = mat_invert(mat_translate(0, 0, 0.8)), tb = mat_invert(mat_translate(0, 0, -0.8));
            S.add_top({model_of(SMCRT_MODEL_INTERSECTION, {prim_sphere(1.0, &ta), prim_sphere(1.0, &tb)}, 0.0), 0.0, 0.0, 0.0, 1.5});
            S.add_top({prim_box(2, 2, 2), 0.0, 0.0, 0.0, 1.0});
        }
    } else if (g == "logo") {
        cfg_fail("need to uncomment inlcude line!");  // setup_logo is disabled in the reference (:328)
    } else {
        cfg_fail("no such routine");  // src/setup.f90:58-59
    }
}

void build_meta(Config& c) {
    std::ostringstream o;
    for (auto& kv : c.dict) {
        bool bare = true;
        for (char ch : kv.first)
            if (!(std::isalnum((unsigned char)ch) || ch == '_' || ch == '-')) bare = false;
        if (bare) o << kv.first;
        else o << '"' << kv.first << '"';
        o << " = " << kv.second << "\n";
    }
    c.meta_text = o.str();
}

Config* load_text(const std::string& text, const std::string& res_dir) {
    Table root = toml_lite::parse(text);
    Config* c = new Config();
    try {
        c->res_dir = res_dir;
        parse_source(root, *c);
        parse_grid(root, *c);
        parse_geometry(root, *c);
        parse_detectors(root, *c);
        parse_output_sim(root, *c);
        build_scene(*c);
        build_meta(*c);
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

bool mkdir_p(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && cur != "/") {
                struct stat st;
                if (stat(cur.c_str(), &st) != 0 && mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) return false;
            }
        }
        if (i < path.size()) cur += path[i];
    }
    return true;
}

}  // namespace

struct smcrt_config {
    Config c;
};

extern "C" {

// host errors are reported through smcrt_last_error() of the engine TU
void smcrt_set_error_(const char* msg);

static int host_fail(const std::string& m) {
    smcrt_set_error_(m.c_str());
    return -1;
}

int smcrt_config_loads(const char* toml_text, const char* res_dir, smcrt_config** out) {
    if (!toml_text || !out) return host_fail("smcrt_config_loads: null argument");
    try {
        Config* c = load_text(toml_text, res_dir ? res_dir : "");
        smcrt_config* h = new smcrt_config{std::move(*c)};
        delete c;
        *out = h;
        return 0;
    } catch (const CfgError& e) {
        return host_fail(e.msg);
    } catch (const std::exception& e) {
        return host_fail(e.what());
    }
}
int smcrt_config_load(const char* toml_path, const char* res_dir, smcrt_config** out) {
    if (!toml_path) return host_fail("smcrt_config_load: null path");
    std::ifstream f(toml_path);
    if (!f) return host_fail(std::string("cannot open ") + toml_path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string rd;
    if (res_dir) rd = res_dir;
    else {
        std::string p = toml_path;
        size_t k = p.find_last_of('/');
        rd = k == std::string::npos ? "." : p.substr(0, k);
    }
    return smcrt_config_loads(ss.str().c_str(), rd.c_str(), out);
}
void smcrt_config_free(smcrt_config* cfg) { delete cfg; }

int smcrt_config_grid(const smcrt_config* cfg, int32_t n[3], double he[3]) {
    n[0] = cfg->c.nxg; n[1] = cfg->c.nyg; n[2] = cfg->c.nzg;
    he[0] = cfg->c.xmax; he[1] = cfg->c.ymax; he[2] = cfg->c.zmax;
    return 0;
}
int64_t smcrt_config_nphotons(const smcrt_config* cfg) { return cfg->c.nphotons; }
int64_t smcrt_config_iseed(const smcrt_config* cfg) { return cfg->c.iseed; }
const char* smcrt_config_geom_name(const smcrt_config* cfg) { return cfg->c.geom.c_str(); }
const char* smcrt_config_source_name(const smcrt_config* cfg) { return cfg->c.source.c_str(); }
int smcrt_config_render_source(const smcrt_config* cfg) { return cfg->c.render_source ? 1 : 0; }
int smcrt_config_source(const smcrt_config* cfg, int32_t* kind, int32_t* subtype, double p[SMCRT_SOURCE_PARAMS]) {
    *kind = cfg->c.src_kind;
    *subtype = cfg->c.src_subtype;
    std::memcpy(p, cfg->c.src, sizeof(double) * SMCRT_SOURCE_PARAMS);
    return 0;
}
int smcrt_config_n_detectors(const smcrt_config* cfg) { return (int)cfg->c.dets.size(); }
int smcrt_config_detectors(const smcrt_config* cfg, int32_t* kind, double* p, int32_t* nbins) {
    for (size_t i = 0; i < cfg->c.dets.size(); ++i) {
        kind[i] = cfg->c.dets[i].kind;
        nbins[i] = cfg->c.dets[i].nbins;
        std::memcpy(p + SMCRT_DET_PARAMS * i, cfg->c.dets[i].p, sizeof(double) * SMCRT_DET_PARAMS);
    }
    return 0;
}
const char* smcrt_config_detector_id(const smcrt_config* cfg, int i) {
    if (i < 0 || i >= (int)cfg->c.dets.size()) return "";
    return cfg->c.dets[i].id.c_str();
}
int smcrt_config_scene_sizes(const smcrt_config* cfg, int32_t* n_nodes, int32_t* n_top) {
    *n_nodes = (int32_t)cfg->c.scene.kind.size();
    *n_top = (int32_t)cfg->c.scene.top_node.size();
    return 0;
}
int smcrt_config_scene(const smcrt_config* cfg, int32_t* kind, int32_t* first_child, int32_t* n_child, double* xform,
                       double* params, int32_t* top_node, double* mus, double* mua, double* hgg, double* n_ref) {
    const FlatScene& S = cfg->c.scene;
    const size_t nn = S.kind.size(), nt = S.top_node.size();
    std::memcpy(kind, S.kind.data(), 4 * nn);
    std::memcpy(first_child, S.first_child.data(), 4 * nn);
    std::memcpy(n_child, S.n_child.data(), 4 * nn);
    std::memcpy(xform, S.xform.data(), 8 * 16 * nn);
    std::memcpy(params, S.params.data(), 8 * SMCRT_NODE_PARAMS * nn);
    std::memcpy(top_node, S.top_node.data(), 4 * nt);
    std::memcpy(mus, S.mus.data(), 8 * nt);
    std::memcpy(mua, S.mua.data(), 8 * nt);
    std::memcpy(hgg, S.hgg.data(), 8 * nt);
    std::memcpy(n_ref, S.n.data(), 8 * nt);
    return 0;
}
const char* smcrt_config_metadata(const smcrt_config* cfg) { return cfg->c.meta_text.c_str(); }

int smcrt_config_apply(const smcrt_config* cfg, smcrt_ctx* ctx) {
    const Config& c = cfg->c;
    const FlatScene& S = c.scene;
    int rc = smcrt_set_grid(ctx, c.nxg, c.nyg, c.nzg, c.xmax, c.ymax, c.zmax);
    if (rc) return rc;
    rc = smcrt_set_scene(ctx, (int)S.kind.size(), S.kind.data(), S.first_child.data(), S.n_child.data(), S.xform.data(),
                         S.params.data(), (int)S.top_node.size(), S.top_node.data(), S.mus.data(), S.mua.data(), S.hgg.data(),
                         S.n.data());
    if (rc) return rc;
    rc = smcrt_set_source(ctx, c.src_kind, c.src_subtype, c.src);
    if (rc) return rc;
    std::vector<int32_t> kind, nb;
    std::vector<double> p;
    for (auto& d : c.dets) {
        kind.push_back(d.kind);
        nb.push_back(d.nbins);
        p.insert(p.end(), d.p, d.p + SMCRT_DET_PARAMS);
    }
    rc = smcrt_set_detectors(ctx, (int)c.dets.size(), kind.data(), p.data(), nb.data());
    if (rc) return rc;
    std::vector<int32_t> track;
    bool any = false;
    for (auto& d : c.dets) { track.push_back(d.track ? 1 : 0); any = any || d.track; }
    return any ? smcrt_set_track_history(ctx, (int)track.size(), track.data()) : 0;
}
int smcrt_config_detector_track(const smcrt_config* cfg, int i) {
    return (cfg && i >= 0 && i < (int)cfg->c.dets.size() && cfg->c.dets[(size_t)i].track) ? 1 : 0;
}
const char* smcrt_config_history_filename(const smcrt_config* cfg) { return cfg ? cfg->c.history_file.c_str() : ""; }

int smcrt_history_write(const char* path, int64_t n, int max_vertices, const float* vertices, const int32_t* counts) {
    if (!path || n < 0 || max_vertices < 1 || (n > 0 && (!vertices || !counts))) return host_fail("smcrt_history_write: invalid arguments");
    const std::string p = path;
    const bool obj = p.find("obj") != std::string::npos, ply = p.find("ply") != std::string::npos, json = p.find("json") != std::string::npos;
    if (!obj && !ply && !json) return host_fail("Unsupported filetype for track History!");  // init_historyStack, :55
    FILE* f = std::fopen(path, "w");
    if (!f) return host_fail(std::string("cannot open ") + path);
    auto cnt = [&](int64_t k) { return std::min<int32_t>(std::max<int32_t>(counts[k], 0), max_vertices); };
    auto v = [&](int64_t k, int j) { return vertices + ((size_t)k * max_vertices + (size_t)j) * 4; };
    int64_t nv = 0, ne = 0;
    for (int64_t k = 0; k < n; ++k) { nv += cnt(k); ne += std::max(cnt(k) - 1, 0); }
    if (obj) {  // obj_writer :184-226 + finish: all "v" lines, then the "l" lines
        for (int64_t k = 0; k < n; ++k)
            for (int j = 0; j < cnt(k); ++j) std::fprintf(f, "v %15.8E %15.8E %15.8E \n", v(k, j)[0], v(k, j)[1], v(k, j)[2]);
        int64_t base = 1;
        for (int64_t k = 0; k < n; ++k) {
            if (cnt(k) >= 2) {
                std::fputs("l ", f);
                for (int j = 0; j < cnt(k); ++j) std::fprintf(f, "%lld ", (long long)(base + j));
                std::fputs("\n", f);
            }
            base += cnt(k);
        }
    } else if (ply) {  // ply_writer :228-273 + the header fix-up of finish
        std::fprintf(f, "ply\nformat ascii 1.0\nelement vertex %lld\nproperty float x\nproperty float y\nproperty float z\nelement edge %lld\n"
                        "property int vertex1\nproperty int vertex2\nend_header\n", (long long)nv, (long long)ne);
        for (int64_t k = 0; k < n; ++k)
            for (int j = 0; j < cnt(k); ++j) std::fprintf(f, "%15.8E %15.8E %15.8E \n", v(k, j)[0], v(k, j)[1], v(k, j)[2]);
        int64_t base = 0;
        for (int64_t k = 0; k < n; ++k) {
            for (int j = 0; j + 1 < cnt(k); ++j) std::fprintf(f, "%lld %lld \n", (long long)(base + j), (long long)(base + j + 1));
            base += cnt(k);
        }
    } else {  // json_writer :275-311 + finish
        std::fputs("{\n", f);
        for (int64_t k = 0; k < n; ++k) {
            std::fprintf(f, "%s\"%lld_0\": [\n", k ? ",\n" : "", (long long)k);
            for (int j = 0; j < cnt(k); ++j)
                std::fprintf(f, "[%15.8E,%15.8E,%15.8E]%s\n", v(k, j)[0], v(k, j)[1], v(k, j)[2], j + 1 < cnt(k) ? "," : "");
            std::fputs("]\n", f);
        }
        std::fputs("}\n", f);
    }
    std::fclose(f);
    return 0;
}

// normalise_fluence, src/writer.f90:25-52.  The factor mixes real32 literals (2._sp) with real64 extents;
// Fortran promotes to real64, and array*factor is real32*real64 -> real64 -> stored real32.
int smcrt_inverse_evaluate(int n_det, const double* totals, const double* targets, int64_t nphotons, double* error) {
    if (n_det < 0 || (n_det > 0 && (!totals || !targets)) || !error || nphotons < 1) return fail("smcrt_inverse_evaluate: invalid arguments");
    double e = 0.0;
    int counter = 0;
    for (int i = 0; i < n_det; ++i)
        if (targets[i] != -1.0) {
            e += std::fabs(totals[i] / (double)nphotons - targets[i]);
            ++counter;
        }
    if (counter == 0) return fail("smcrt_inverse_evaluate: no detector has a target value");
    *error = -e / counter;
    return 0;
}
int smcrt_escape_cell_centre(int m, int n, int o, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax,
                             const double* rot_z, const double* rot_off, const double* grid_pos, double* out_xyz) {
    if (!out_xyz || nxg < 1 || nyg < 1 || nzg < 1) return fail("smcrt_escape_cell_centre: invalid arguments");
    double v[3] = {(((double)m - 0.5) / nxg) * 2.0 * xmax - xmax, (((double)n - 0.5) / nyg) * 2.0 * ymax - ymax,
                   (((double)o - 0.5) / nzg) * 2.0 * zmax - zmax};
    for (const double* M : {rot_z, rot_off})
        if (M) {  // vec .dot. mat (src/vector_class.f90:292-304): row vector times the matrix, translation in row 4
            double r[3];
            for (int j = 0; j < 3; ++j) r[j] = v[0] * M[0 + 4 * j] + v[1] * M[1 + 4 * j] + v[2] * M[2 + 4 * j] + M[3 + 4 * j];
            v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
        }
    for (int a = 0; a < 3; ++a) out_xyz[a] = v[a] + (grid_pos ? grid_pos[a] : 0.0);
    return 0;
}
int smcrt_normalise_fluence(float* array, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax, int64_t nphotons) {
    const double num = (2.0 * xmax * 2.0 * ymax * 2.0 * zmax);
    // (the reference's nphotons is a default integer; this engine takes 64-bit counts and traces 4e9 packets in a second, so the
    // count is NOT narrowed: a 3e9-packet job would otherwise normalise by a negative number)
    if (nphotons <= 0) return host_fail("smcrt_normalise_fluence: nphotons must be positive");
    const double den = ((double)nphotons * (2.0 * xmax / nxg) * (2.0 * ymax / nyg) * (2.0 * zmax / nzg));
    const double f = num / den;
    const size_t n = (size_t)nxg * nyg * nzg;
    for (size_t i = 0; i < n; ++i) array[i] = (float)((double)array[i] * f);
    return 0;
}

// write_3d_r4_nrrd + write_hdr, src/writer.f90:304-337, 382-424
int smcrt_write_nrrd_f32(const char* path, const float* data, int nxg, int nyg, int nzg, const char* meta) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return host_fail(std::string("cannot open ") + path);
    // sizes are written REVERSED (str(sizes(3)) first) exactly like write_hdr
    std::fprintf(f, "NRRD0004\ntype: float\ndimension: 3\nsizes: %d %d %d\nspace dimension: 3\nencoding: raw\nendian: little\n",
                 nzg, nyg, nxg);
    if (meta && *meta) std::fputs(meta, f);
    std::fputs("\n\n", f);  // write(u,"(A)") new_line("C")  -> the newline character followed by the record end
    const size_t n = (size_t)nxg * nyg * nzg;
    size_t w = std::fwrite(data, sizeof(float), n, f);
    std::fclose(f);
    return w == n ? 0 : host_fail("short write");
}

// write_detected_photons, src/writer.f90:55-134: a raw stream of real64
int smcrt_write_detectors(const smcrt_config* cfg, const double* det_bins, const char* out_dir) {
    const Config& c = cfg->c;
    if (!mkdir_p(out_dir)) return host_fail(std::string("cannot create ") + out_dir);
    size_t off = 0;
    for (size_t i = 0; i < c.dets.size(); ++i) {
        const DetCfg& d = c.dets[i];
        const int stored = d.nbins + 1;
        const size_t count = d.kind == SMCRT_DET_CAMERA ? (size_t)stored * stored : (size_t)stored;
        std::string path = std::string(out_dir) + "/detector_" + std::to_string(i + 1) + ".dat";
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) return host_fail("cannot open " + path);
        auto put = [&](double v) { std::fwrite(&v, 8, 1, f); };
        auto put_id = [&]() {
            put((double)d.id.size());
            for (char ch : d.id) put((double)(unsigned char)ch);
        };
        double bw = 1.0;
        if (d.kind == SMCRT_DET_CIRCLE) {
            bw = d.nbins == 0 ? 1.0 : d.p[6] / d.nbins;
            put(1.0); put_id(); put((double)c.nphotons); put(d.p[6]);
            for (int k = 0; k < 6; ++k) put(d.p[k]);
            for (int j = 1; j <= stored; ++j) { put((j - 0.5) * bw); put(det_bins[off + j - 1]); }
        } else if (d.kind == SMCRT_DET_FIBRE) {
            bw = d.nbins == 0 ? 1.0 : d.p[16] / 2 / d.nbins;
            put(2.0); put_id(); put((double)c.nphotons);
            for (int k = 0; k < 6; ++k) put(d.p[k]);
            for (int k = 6; k <= 16; ++k) put(d.p[k]);
            for (int j = 1; j <= stored; ++j) { put((j - 0.5) * bw); put(det_bins[off + j - 1]); }
        } else if (d.kind == SMCRT_DET_ANNULUS) {
            bw = d.nbins == 0 ? 1.0 : (d.p[7] - d.p[6]) / d.nbins;
            put(3.0); put_id(); put((double)c.nphotons); put(d.p[6]); put(d.p[7]);
            for (int k = 0; k < 6; ++k) put(d.p[k]);
            for (int j = 1; j <= stored; ++j) { put((j - 0.5) * bw + d.p[6]); put(det_bins[off + j - 1]); }
        }  // camera: "not yet implmented" in the reference — an empty file is created, like `open(status='REPLACE')`
        std::fclose(f);
        off += count;
    }
    return 0;
}

// checkpoint, src/writer.f90:426-457 (two formatted lines, then the stream-appended jmean)
int smcrt_checkpoint_write(const char* path, const char* toml_filename, int64_t nphotons_run, const float* jmean, int64_t n_voxels) {
    if (!path || !toml_filename || !jmean || n_voxels < 0) return host_fail("smcrt_checkpoint_write: invalid arguments");
    const std::string tmp = std::string(path) + ".tmp";  // written aside and renamed: a kill mid-write leaves the previous checkpoint
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return host_fail(std::string("cannot write checkpoint ") + path);
    std::fprintf(f, "tomlfile=%s\nphotons_run=%lld\n", toml_filename, (long long)nphotons_run);
    const size_t w = std::fwrite(jmean, sizeof(float), (size_t)n_voxels, f);
    const bool ok = w == (size_t)n_voxels && std::fclose(f) == 0;
    if (!ok || std::rename(tmp.c_str(), path) != 0) return host_fail(std::string("cannot write checkpoint ") + path);
    return 0;
}
int smcrt_checkpoint_read(const char* path, char* toml_out, int toml_cap, int64_t* nphotons_run, float* jmean, int64_t n_voxels) {
    if (!path) return host_fail("smcrt_checkpoint_read: invalid arguments");
    FILE* f = std::fopen(path, "rb");
    if (!f) return host_fail(std::string("cannot open checkpoint ") + path);
    auto line = [&](std::string& out) {
        out.clear();
        int ch;
        while ((ch = std::fgetc(f)) != EOF && ch != '\n') out.push_back((char)ch);
        return ch != EOF;
    };
    std::string l1, l2;
    if (!line(l1) || !line(l2) || l1.find('=') == std::string::npos || l2.find('=') == std::string::npos) {
        std::fclose(f);
        return host_fail(std::string("malformed checkpoint header in ") + path);
    }
    const std::string name = l1.substr(l1.find('=') + 1);  // kernelsMod.f90:55-57: everything after the first '='
    if (toml_out && toml_cap > 0) std::snprintf(toml_out, (size_t)toml_cap, "%s", name.c_str());
    if (nphotons_run) *nphotons_run = std::atoll(l2.substr(l2.find('=') + 1).c_str());
    int rc = 0;
    if (jmean && std::fread(jmean, sizeof(float), (size_t)n_voxels, f) != (size_t)n_voxels)
        rc = host_fail(std::string("checkpoint ") + path + " holds fewer voxels than the grid of its input deck");
    std::fclose(f);
    return rc;
}

// default_MCRT, src/kernelsMod.f90:29-83
int smcrt_default_mcrt(const char* toml_path, const char* res_dir, const char* out_dir, int n_gpus, int tally_mode,
                       int survival_bias, int64_t nphotons, double* photons_per_s, smcrt_counters* counters) {
    smcrt_config* cfg = nullptr;
    int rc = smcrt_config_load(toml_path, res_dir, &cfg);
    if (rc) return rc;
    // load_checkpoint (kernelsMod.f90:52-72): the checkpoint names the input deck it belongs to; that deck is the one that runs
    int64_t photons_done = 0;
    std::string ckpt_in;
    if (cfg->c.loadckpt) {
        ckpt_in = cfg->c.ckptfile;
        char name[1024];
        if ((rc = smcrt_checkpoint_read(ckpt_in.c_str(), name, (int)sizeof name, &photons_done, nullptr, 0))) { smcrt_config_free(cfg); return rc; }
        smcrt_config_free(cfg);
        cfg = nullptr;
        if ((rc = smcrt_config_load(name, res_dir, &cfg))) return rc;
        toml_path = nullptr;  // (the deck's own name is used for later checkpoints)
        cfg->c.ckpt_deck = name;
    } else
        cfg->c.ckpt_deck = toml_path ? toml_path : "";
    Config& c = cfg->c;
    if (nphotons > 0) c.nphotons = nphotons;
    smcrt_ctx* ctx = nullptr;
    rc = smcrt_create(&ctx, n_gpus, nullptr);
    if (rc) { smcrt_config_free(cfg); return rc; }
    auto cleanup = [&](int r) {
        smcrt_destroy(ctx);
        smcrt_config_free(cfg);
        return r;
    };
    if ((rc = smcrt_config_apply(cfg, ctx))) return cleanup(rc);
    if (tally_mode < 0) tally_mode = SMCRT_TALLY_ABSORB | (c.render_source ? SMCRT_TALLY_EMISSION : 0);
    const size_t nv = (size_t)c.nxg * c.nyg * c.nzg;
    std::vector<float> jmean(nv, 0.f), absorb(nv, 0.f), emission(nv, 0.f);
    if (!ckpt_in.empty()) {  // resume: jmean of the packets already run (the only tally the reference checkpoints)
        if ((rc = smcrt_checkpoint_read(ckpt_in.c_str(), nullptr, 0, nullptr, jmean.data(), (int64_t)nv))) return cleanup(rc);
        if (photons_done < 0 || photons_done > c.nphotons) return cleanup(host_fail("checkpoint has run more packets than the deck asks for"));
    }
    // run_MCRT's loop (kernelsMod.f90:1861-1888) with its `mod(j, ckptfreq) == 0 -> checkpoint` cut points.  Packet streams depend
    // on (seed, packet id) only, so the pieces -- and a resumed run, which starts at id photons_done -- are ONE job.
    const bool ckpt_on = (tally_mode & SMCRT_TALLY_PATHLENGTH) && c.ckptfreq > 0 && c.ckptfreq < c.nphotons;
    auto t0 = std::chrono::steady_clock::now();
    auto last_write = t0;
    int64_t done = photons_done;
    std::vector<float> snap;
    while (done < c.nphotons) {
        int64_t piece = c.nphotons - done;
        if (ckpt_on) {  // up to the next multiple of ckptfreq that is >= 64 Mi packets away (a piece shorter than that is all overhead)
            static const char* mp = std::getenv("SMCRT_CKPT_MIN_PIECE");  // tests
            const int64_t min_piece = mp ? std::max<int64_t>(1, std::atoll(mp)) : (64ll << 20);
            const int64_t k = std::max<int64_t>(1, (min_piece + c.ckptfreq - 1) / c.ckptfreq);
            piece = std::min<int64_t>(piece, (done / c.ckptfreq + k) * c.ckptfreq - done);
        }
        if ((rc = smcrt_run(ctx, piece, (uint64_t)c.iseed, done, tally_mode, survival_bias, -1.0, -1.0))) return cleanup(rc);
        done += piece;
        const auto now = std::chrono::steady_clock::now();
        static const char* ms = std::getenv("SMCRT_CKPT_MIN_SECONDS");  // tests
        if (ckpt_on && done < c.nphotons && std::chrono::duration<double>(now - last_write).count() >= (ms ? std::atof(ms) : 2.0)) {
            snap = jmean;  // what the checkpoint held + what the device has accumulated since
            if ((rc = smcrt_fetch(ctx, snap.data(), nullptr, nullptr, nullptr, nullptr, 1))) return cleanup(rc);
            if ((rc = smcrt_checkpoint_write(c.ckptfile.c_str(), c.ckpt_deck.c_str(), done, snap.data(), (int64_t)nv))) return cleanup(rc);
            last_write = std::chrono::steady_clock::now();
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    const double secs = std::chrono::duration<double>(t1 - t0).count();
    if (photons_per_s) *photons_per_s = (double)(c.nphotons - photons_done) / secs;  // print*,"Photons/s: ", kernelsMod.f90:1897
    std::vector<double> bins((size_t)std::max<int64_t>(1, smcrt_det_bins_total(ctx)), 0.0);
    smcrt_counters cn{};
    if ((rc = smcrt_fetch(ctx, jmean.data(), absorb.data(), emission.data(), bins.data(), &cn, 1))) return cleanup(rc);
    if (counters) *counters = cn;
    // A resumed run restores jmean from the checkpoint, but absorb, emission, the detector bins and the scatter count cover only the
    // packets traced NOW (ids photons_done .. nphotons).  The reference keeps them consistent by reducing state%nphotons by the
    // packets already run (kernelsMod.f90:72); the same count is used here for everything but jmean, which holds the whole job.
    const int64_t n_total = c.nphotons, n_traced = c.nphotons - photons_done;
    std::printf(" Average # of scatters per photon: %.10g\n", cn.nscatt / (double)std::max<int64_t>(n_traced, 1));
    // finalise: metadata then files (kernelsMod.f90:2376-2392)
    {
        char b[128];
        c.dict_sets("grid_data", "fluence map");
        std::snprintf(b, sizeof b, "%.7f %.7f %.7f", c.xmax, c.ymax, c.zmax);
        c.dict_sets("real_size", b);
        c.dict_seti("nphotons", n_traced);
        c.dict_sets("source", c.source);
        c.dict_sets("experiment", c.geom);
        build_meta(c);
    }
    const std::string od = out_dir ? out_dir : "data";
    for (const char* sub : {"/jmean", "/emission", "/absorb", "/detectors"})
        if (!mkdir_p(od + sub)) return cleanup(host_fail("cannot create output directory " + od + sub));
    if (tally_mode & SMCRT_TALLY_PATHLENGTH) {
        smcrt_normalise_fluence(jmean.data(), c.nxg, c.nyg, c.nzg, c.xmax, c.ymax, c.zmax, n_total);
        if ((rc = smcrt_write_nrrd_f32((od + "/jmean/" + c.outfile).c_str(), jmean.data(), c.nxg, c.nyg, c.nzg, c.meta_text.c_str())))
            return cleanup(rc);
    }
    c.nphotons = std::max<int64_t>(n_traced, 1);  // emission, detector-file headers: the packets traced by this invocation
    smcrt_normalise_fluence(emission.data(), c.nxg, c.nyg, c.nzg, c.xmax, c.ymax, c.zmax, c.nphotons);
    if ((rc = smcrt_write_nrrd_f32((od + "/emission/" + c.rendersourcefile).c_str(), emission.data(), c.nxg, c.nyg, c.nzg, c.meta_text.c_str())))
        return cleanup(rc);
    if ((rc = smcrt_write_nrrd_f32((od + "/absorb/absorb.nrrd").c_str(), absorb.data(), c.nxg, c.nyg, c.nzg, c.meta_text.c_str())))
        return cleanup(rc);
    if (!c.dets.empty())
        if ((rc = smcrt_write_detectors(cfg, bins.data(), (od + "/detectors").c_str()))) return cleanup(rc);
    bool any_track = false;
    for (auto& d : c.dets) any_track = any_track || d.track;
    if (any_track) {  // history%write for every packet that hit a tracking detector; "<name>_000.<ext>" like init_historyStack (:44-45)
        int64_t total = 0;
        if ((rc = smcrt_history_hits(ctx, 0, nullptr, nullptr, &total))) return cleanup(rc);
        const int64_t keep = std::min<int64_t>(total, 100000);  // (a file of polylines: a diagnostic, not a tally)
        std::vector<uint64_t> ids((size_t)std::max<int64_t>(keep, 1));
        std::vector<int32_t> det((size_t)std::max<int64_t>(keep, 1));
        if ((rc = smcrt_history_hits(ctx, keep, ids.data(), det.data(), &total))) return cleanup(rc);
        const int maxv = 256;
        std::vector<float> verts((size_t)std::max<int64_t>(keep, 1) * maxv * 4);
        std::vector<int32_t> nvert((size_t)std::max<int64_t>(keep, 1), 0), hit((size_t)std::max<int64_t>(keep, 1), 0);
        if (keep > 0 && (rc = smcrt_history_replay(ctx, keep, ids.data(), (uint64_t)c.iseed, survival_bias, maxv, verts.data(), nvert.data(), hit.data())))
            return cleanup(rc);
        for (int64_t k = 0; k < keep; ++k) nvert[(size_t)k] = hit[(size_t)k] > 0 ? hit[(size_t)k] : 0;  // the list as it stood at the hit
        std::string name = c.history_file;
        const size_t dot = name.find('.');
        if (dot != std::string::npos) name = name.substr(0, dot) + "_000" + name.substr(dot);
        if ((rc = smcrt_history_write((od + "/" + name).c_str(), keep, maxv, verts.data(), nvert.data()))) return cleanup(rc);
    }
    return cleanup(0);
}

}  // extern "C"
