// f16_tables_host.h -- host-side table loading (see f16_tables_host.cpp)
#pragma once
#include <stddef.h>

#include <string>
#include <vector>

namespace f16 {
// path: blob file or directory of C/*.dat; nullptr/"" = search ($F16_TABLE_PATH, <lib_dir>/../data, ./C, lib_dir)
bool load_canonical(const char* path, const std::string& lib_dir, std::vector<double>& payload, std::string& source,
                    std::string& err);
bool check_grids(const std::vector<double>& payload, std::string& err);
void payload_sha256_hex(const std::vector<double>& payload, char out65[65]);
size_t canon_table_offset(int table_id);
void build_hifi_image(const std::vector<double>& payload, bool clr_from_file, std::vector<double>& img);
void build_hifi_fast_image(const std::vector<double>& payload, bool clr_from_file, std::vector<double>& img);
void build_lofi_image(std::vector<double>& img);
}  // namespace f16
