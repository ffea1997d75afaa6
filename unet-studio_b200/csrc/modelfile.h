// Model file (.nz) codec and model-level load / save (modelfile.cpp).
#pragma once
#include <string>
#include <vector>

#include "model.h"

namespace u3d {
struct NzFile;
int nz_add(NzFile& f, const std::string& name, int type, int rows, int cols, const void* data);
int nz_save(const NzFile& f, const std::string& path);
int nz_load(const std::string& path, NzFile& f);
int nz_read_f32(const NzFile& f, const std::string& name, std::vector<float>& out);
int model_to_nz(Model& m, NzFile& f);
int nz_model_header(const NzFile& f, int& in_c, int& out_c, std::string& architecture);
int nz_to_model(const NzFile& f, Model& m);
int model_momentum_to_nz(Model& m, NzFile& f);
int nz_to_model_momentum(const NzFile& f, Model& m);
int model_export_raw(Model& m, const std::string& dir);
NzFile* nz_new();
void nz_delete(NzFile* f);
int nz_count(const NzFile& f);
int nz_info(const NzFile& f, int i, std::string& name, int& type, int& rows, int& cols);
}  // namespace u3d
