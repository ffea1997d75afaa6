// Model file (.nz) codec and the model-level load / save built on it.
//
// Reference: load_from_file / save_to_file, /root/reference/main.cpp:157-233.  The file is a gzip stream holding MATLAB Level-4
// MAT matrices (TIPL tipl::io::gz_mat_read / gz_mat_write -- TIPL is not vendored, so its dialect is PARITY UNPINNED; what is pinned
// here is the public Level-4 container, checked against scipy.io in tests/test_modelfile_cpu.py):
//     per matrix: int32 type, mrows, ncols, imagf, namlen (incl. NUL); name; mrows*ncols elements, column-major
//     type = 1000*M + 100*O + 10*P + T with M = 0 (little endian), P: 0 f64, 1 f32, 2 i32, 3 i16, 4 u16, 5 u8; T: 0 numeric, 1 text
// Logical layout written / expected (main.cpp:212-231): channels (i32[2]), architecture (text), dimension (3), voxel_size (3),
// fov_strategy, preproc, orientation, postproc (text), training_errors / testing_errors (3 x steps: ce, dice, mse),
// single_component_label, tensor{i} = parameters()[i] as rows = numel/size(0), cols = size(0) in native contiguous element order.
// The reference writes tensor{i} with tipl::io::sloped (apply_slope, min_size_for_mask_slope = 1024): an integer-quantised matrix
// with companion slope / intercept matrices.  That encoding lives in TIPL; this reader ASSUMES companions named "<name>.slope" and
// "<name>.inter" (scalar or one per column, value = stored*slope + inter) and otherwise converts any numeric type to fp32; this
// writer stores fp32 (type 10) exactly, which read_as_type<float> accepts.
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "modelfile.h"

namespace u3d {

struct NzMatrix {
    std::string name;
    int type = 10;
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;
    size_t count() const { return size_t(rows) * cols; }
};

struct NzFile {
    std::vector<NzMatrix> mats;
    const NzMatrix* find(const std::string& name) const {
        for (const auto& m : mats)
            if (m.name == name) return &m;
        return nullptr;
    }
};

NzFile* nz_new() { return new NzFile; }
void nz_delete(NzFile* f) { delete f; }
int nz_count(const NzFile& f) { return int(f.mats.size()); }
int nz_info(const NzFile& f, int i, std::string& name, int& type, int& rows, int& cols) {
    if (i < 0 || i >= int(f.mats.size())) { set_error("matrix index out of range"); return 1; }
    name = f.mats[size_t(i)].name; type = f.mats[size_t(i)].type; rows = f.mats[size_t(i)].rows; cols = f.mats[size_t(i)].cols;
    return 0;
}

static int elem_size(int type) {
    switch ((type / 10) % 10) {
        case 0: return 8;
        case 1: return 4;
        case 2: return 4;
        case 3: return 2;
        case 4: return 2;
        case 5: return 1;
        default: return 0;
    }
}

static double elem_value(const NzMatrix& m, size_t i) {
    const uint8_t* p = m.data.data() + i * elem_size(m.type);
    switch ((m.type / 10) % 10) {
        case 0: { double v; std::memcpy(&v, p, 8); return v; }
        case 1: { float v; std::memcpy(&v, p, 4); return v; }
        case 2: { int32_t v; std::memcpy(&v, p, 4); return v; }
        case 3: { int16_t v; std::memcpy(&v, p, 2); return v; }
        case 4: { uint16_t v; std::memcpy(&v, p, 2); return v; }
        default: return *p;
    }
}

int nz_add(NzFile& f, const std::string& name, int type, int rows, int cols, const void* data) {
    const int es = elem_size(type);
    if (es == 0 || rows < 0 || cols < 0 || type < 0 || type >= 1000 || (type / 100) % 10 != 0) { set_error("nz: unsupported matrix type " + std::to_string(type)); return 1; }
    if (name.empty()) { set_error("nz: empty matrix name"); return 1; }
    NzMatrix m;
    m.name = name; m.type = type; m.rows = rows; m.cols = cols;
    m.data.resize(m.count() * es);
    if (!m.data.empty()) std::memcpy(m.data.data(), data, m.data.size());
    for (auto& old : f.mats)
        if (old.name == name) { old = std::move(m); return 0; }
    f.mats.push_back(std::move(m));
    return 0;
}

static void nz_add_text(NzFile& f, const std::string& name, const std::string& text) { nz_add(f, name, 51, 1, int(text.size()), text.data()); }

int nz_save(const NzFile& f, const std::string& path) {
    gzFile gz = gzopen(path.c_str(), "wb6");
    if (!gz) { set_error("cannot open " + path + " for writing"); return 1; }
    bool ok = true;
    for (const auto& m : f.mats) {
        const int32_t hdr[5] = {m.type, m.rows, m.cols, 0, int32_t(m.name.size() + 1)};
        ok = ok && gzwrite(gz, hdr, sizeof(hdr)) == int(sizeof(hdr));
        ok = ok && gzwrite(gz, m.name.c_str(), unsigned(m.name.size() + 1)) == int(m.name.size() + 1);
        size_t off = 0;
        while (ok && off < m.data.size()) {   // gzwrite takes an unsigned length
            const unsigned n = unsigned(std::min<size_t>(m.data.size() - off, size_t(1) << 30));
            ok = gzwrite(gz, m.data.data() + off, n) == int(n);
            off += n;
        }
    }
    if (gzclose(gz) != Z_OK) ok = false;
    if (!ok) { set_error("write error on " + path); return 1; }
    return 0;
}

int nz_load(const std::string& path, NzFile& f) {
    gzFile gz = gzopen(path.c_str(), "rb");   // also reads an uncompressed MAT-v4 file (zlib's transparent mode)
    if (!gz) { set_error("cannot open " + path); return 1; }
    f.mats.clear();
    std::string err;
    for (;;) {
        int32_t hdr[5];
        const int got = gzread(gz, hdr, sizeof(hdr));
        if (got == 0) break;
        if (got != int(sizeof(hdr))) { err = "truncated matrix header"; break; }
        NzMatrix m;
        m.type = hdr[0]; m.rows = hdr[1]; m.cols = hdr[2];
        if (m.type < 0 || m.type >= 1000 || (m.type / 100) % 10 != 0 || elem_size(m.type) == 0 || m.rows < 0 || m.cols < 0 || hdr[3] != 0 ||
            hdr[4] <= 0 || hdr[4] > 4096) { err = "not a little-endian Level-4 MAT matrix (type " + std::to_string(hdr[0]) + ")"; break; }
        std::string name(size_t(hdr[4]), '\0');
        if (gzread(gz, &name[0], unsigned(hdr[4])) != hdr[4]) { err = "truncated matrix name"; break; }
        name.resize(std::strlen(name.c_str()));
        m.name = name;
        const size_t bytes = m.count() * elem_size(m.type);
        if (bytes > (size_t(1) << 34)) { err = "matrix " + name + " is implausibly large"; break; }
        m.data.resize(bytes);
        size_t off = 0;
        while (off < bytes) {
            const unsigned n = unsigned(std::min<size_t>(bytes - off, size_t(1) << 30));
            if (gzread(gz, m.data.data() + off, n) != int(n)) { err = "truncated data of matrix " + name; break; }
            off += n;
        }
        if (!err.empty()) break;
        f.mats.push_back(std::move(m));
    }
    gzclose(gz);
    if (!err.empty()) { set_error(path + ": " + err); return 1; }
    if (f.mats.empty()) { set_error(path + ": no matrices"); return 1; }
    return 0;
}

// any numeric type -> fp32, with the assumed sloped companions applied
int nz_read_f32(const NzFile& f, const std::string& name, std::vector<float>& out) {
    const NzMatrix* m = f.find(name);
    if (!m) { set_error("matrix " + name + " not found"); return 1; }
    out.resize(m->count());
    if ((m->type / 10) % 10 == 1) std::memcpy(out.data(), m->data.data(), out.size() * 4);
    else
        for (size_t i = 0; i < out.size(); ++i) out[i] = float(elem_value(*m, i));
    const NzMatrix* slope = f.find(name + ".slope");
    const NzMatrix* inter = f.find(name + ".inter");
    if (slope || inter) {
        auto coef = [&](const NzMatrix* c, size_t col, double dflt) {
            if (!c || c->count() == 0) return dflt;
            return elem_value(*c, c->count() == size_t(m->cols) ? col : 0);
        };
        for (int c = 0; c < m->cols; ++c) {
            const double s = coef(slope, size_t(c), 1.0), b = coef(inter, size_t(c), 0.0);
            for (int r = 0; r < m->rows; ++r) {
                float& v = out[size_t(c) * m->rows + r];
                v = float(double(v) * s + b);
            }
        }
    }
    return 0;
}

static bool nz_read_text(const NzFile& f, const std::string& name, std::string& out) {
    const NzMatrix* m = f.find(name);
    if (!m) return false;
    out.clear();
    for (size_t i = 0; i < m->count(); ++i) {
        const int c = int(elem_value(*m, i));
        if (c == 0) break;
        out.push_back(char(c));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------------------------
// model level
// ------------------------------------------------------------------------------------------------------------------------------
int model_to_nz(Model& m, NzFile& f) {
    const int32_t ch[2] = {m.in_count, m.out_count};
    nz_add(f, "channels", 20, 1, 2, ch);
    nz_add_text(f, "architecture", m.architecture);
    const int32_t dim[3] = {m.dim[0], m.dim[1], m.dim[2]};
    nz_add(f, "dimension", 20, 1, 3, dim);
    nz_add(f, "voxel_size", 10, 1, 3, m.voxel_size);
    nz_add_text(f, "fov_strategy", m.fov_strategy);
    nz_add_text(f, "preproc", m.preproc);
    nz_add_text(f, "orientation", m.orientation);
    nz_add_text(f, "postproc", m.postproc);
    nz_add(f, "training_errors", 10, 3, int(m.training_errors.size() / 3), m.training_errors.data());
    nz_add(f, "testing_errors", 10, 3, int(m.testing_errors.size() / 3), m.testing_errors.data());
    if (!m.single_component_label.empty())
        nz_add(f, "single_component_label", 20, 1, int(m.single_component_label.size()), m.single_component_label.data());
    std::vector<float> host;
    for (size_t i = 0; i < m.params.size(); ++i) {
        const ParamInfo& p = m.params[i];
        host.resize(size_t(p.numel));
        if (m.get_flat(m.params_base(), int(i), host.data(), 1.f)) return 1;
        const int cols = int(p.shape[0]);
        nz_add(f, "tensor" + std::to_string(i), 10, int(p.numel / cols), cols, host.data());
    }
    return 0;
}

// the structural part of load_from_file (main.cpp:163-190): channels + architecture -> constructor arguments
int nz_model_header(const NzFile& f, int& in_c, int& out_c, std::string& architecture) {
    std::vector<float> ch;
    if (!f.find("channels") || nz_read_f32(f, "channels", ch) || ch.size() < 2 || !nz_read_text(f, "architecture", architecture)) {
        set_error("invalid format");
        return 1;
    }
    in_c = int(ch[0]); out_c = int(ch[1]);
    return 0;
}

int nz_to_model(const NzFile& f, Model& m) {
    std::vector<float> v;
    if (!f.find("dimension") || nz_read_f32(f, "dimension", v) || v.size() < 3) { set_error("invalid format"); return 1; }
    if (m.set_dim(int(v[0]), int(v[1]), int(v[2]))) return 1;
    if (!f.find("voxel_size") || nz_read_f32(f, "voxel_size", v) || v.size() < 3) { set_error("invalid format"); return 1; }
    for (int k = 0; k < 3; ++k) m.voxel_size[k] = v[k];
    nz_read_text(f, "fov_strategy", m.fov_strategy);
    nz_read_text(f, "preproc", m.preproc);
    nz_read_text(f, "orientation", m.orientation);
    nz_read_text(f, "postproc", m.postproc);
    m.single_component_label.clear();
    if (f.find("single_component_label") && !nz_read_f32(f, "single_component_label", v))
        for (float x : v) m.single_component_label.push_back(int(x));
    m.testing_errors.clear();
    m.training_errors.clear();
    if (f.find("testing_errors")) nz_read_f32(f, "testing_errors", m.testing_errors);
    if (f.find("training_errors")) nz_read_f32(f, "training_errors", m.training_errors);
    m.training_errors.resize(m.testing_errors.size());   // main.cpp:191
    for (size_t i = 0; i < m.params.size(); ++i) {
        const std::string name = "tensor" + std::to_string(i);
        const NzMatrix* t = f.find(name);
        if (!t || (nz_read_f32(f, name, v), (long long)v.size() != m.params[i].numel)) {
            set_error("tensor size mismatch at " + name + " " + std::to_string(t ? t->count() : 0) + " not the expected of size " +
                      std::to_string(m.params[i].numel));   // main.cpp:199-201
            return 1;
        }
        if (m.set_param(int(i), v.data())) return 1;
    }
    return 0;
}

// optimizer state next to the model file (train.cpp:787, 945-957 use torch::save / torch::load of the SGD object: a libtorch pickle
// archive that cannot be produced without libtorch).  Same information in the .nz container: momentum{i} per parameter + step state.
int model_momentum_to_nz(Model& m, NzFile& f) {
    std::vector<float> host;
    for (size_t i = 0; i < m.params.size(); ++i) {
        const ParamInfo& p = m.params[i];
        host.resize(size_t(p.numel));
        if (m.get_flat(m.momentum_base(), int(i), host.data(), 1.f)) return 1;
        const int cols = int(p.shape[0]);
        nz_add(f, "momentum" + std::to_string(i), 10, int(p.numel / cols), cols, host.data());
    }
    const float st[2] = {m.lr0, m.mom_initialized ? 1.f : 0.f};
    nz_add(f, "sgd_state", 10, 1, 2, st);
    return 0;
}

int nz_to_model_momentum(const NzFile& f, Model& m) {
    std::vector<float> v;
    for (size_t i = 0; i < m.params.size(); ++i) {
        const std::string name = "momentum" + std::to_string(i);
        if (!f.find(name) || nz_read_f32(f, name, v) || (long long)v.size() != m.params[i].numel) {
            set_error("optimizer file: size mismatch at " + name);
            return 1;
        }
        if (m.set_momentum(int(i), v.data())) return 1;
    }
    if (f.find("sgd_state") && !nz_read_f32(f, "sgd_state", v) && v.size() >= 2) m.mom_initialized = v[1] != 0.f;
    return 0;
}

// raw export of the same logical layout: <dir>/tensor{i}.bin (fp32, native order) + <dir>/model.json
static std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); }
        else if (c == '\n') o += "\\n";
        else if (c == '\r') o += "\\r";
        else if (c == '\t') o += "\\t";
        else o.push_back(c);
    }
    return o;
}

int model_export_raw(Model& m, const std::string& dir) {
    std::string j = "{\n \"channels\": [" + std::to_string(m.in_count) + ", " + std::to_string(m.out_count) + "],\n";
    j += " \"architecture\": \"" + json_escape(m.architecture) + "\",\n";
    j += " \"dimension\": [" + std::to_string(m.dim[0]) + ", " + std::to_string(m.dim[1]) + ", " + std::to_string(m.dim[2]) + "],\n";
    char buf[128];
    std::snprintf(buf, sizeof buf, " \"voxel_size\": [%.9g, %.9g, %.9g],\n", m.voxel_size[0], m.voxel_size[1], m.voxel_size[2]);
    j += buf;
    j += " \"fov_strategy\": \"" + json_escape(m.fov_strategy) + "\", \"preproc\": \"" + json_escape(m.preproc) + "\", \"orientation\": \"" +
         json_escape(m.orientation) + "\", \"postproc\": \"" + json_escape(m.postproc) + "\",\n \"tensors\": [";
    std::vector<float> host;
    for (size_t i = 0; i < m.params.size(); ++i) {
        const ParamInfo& p = m.params[i];
        host.resize(size_t(p.numel));
        if (m.get_flat(m.params_base(), int(i), host.data(), 1.f)) return 1;
        const std::string fn = dir + "/tensor" + std::to_string(i) + ".bin";
        FILE* fp = std::fopen(fn.c_str(), "wb");
        if (!fp || std::fwrite(host.data(), 4, host.size(), fp) != host.size()) {
            if (fp) std::fclose(fp);
            set_error("cannot write " + fn);
            return 1;
        }
        std::fclose(fp);
        j += std::string(i ? "," : "") + "\n  {\"file\": \"tensor" + std::to_string(i) + ".bin\", \"name\": \"" + p.name + "\", \"rows\": " +
             std::to_string(p.numel / p.shape[0]) + ", \"cols\": " + std::to_string(p.shape[0]) + ", \"shape\": [";
        for (size_t k = 0; k < p.shape.size(); ++k) j += std::string(k ? ", " : "") + std::to_string(p.shape[k]);
        j += "]}";
    }
    j += "]\n}\n";
    const std::string fn = dir + "/model.json";
    FILE* fp = std::fopen(fn.c_str(), "wb");
    if (!fp || std::fwrite(j.data(), 1, j.size(), fp) != j.size()) {
        if (fp) std::fclose(fp);
        set_error("cannot write " + fn);
        return 1;
    }
    std::fclose(fp);
    return 0;
}

}  // namespace u3d
