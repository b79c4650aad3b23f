/* The two file formats the replaced blocks read at construction, and the output format names, shared by the
 * reference-facing adapters (B200OfdmChain, B200Blocks).  Same acceptance rules and messages as the reference's
 * loaders: FIRFilter::load_filter_taps (src/FIRFilter.cpp:95-141), MemlessPoly::load_coefficients
 * (src/MemlessPoly.cpp:145-235), FormatConverter (src/FormatConverter.cpp:44, 190-205). */
#pragma once

#include <fstream>
#include <istream>
#include <stdexcept>
#include <string>
#include <vector>

#include "dabmod_b200.h"

namespace b200files {

inline std::vector<float> load_taps(const std::string& file)
{
    std::vector<float> taps;
    if (file == "default") {
        taps.resize(dabmod_b200_default_fir_taps(nullptr, 0));
        dabmod_b200_default_fir_taps(taps.data(), (int)taps.size());
        return taps;
    }
    std::ifstream in(file);
    if (!in) throw std::runtime_error("FIRFilter: Could not open file with taps! " + file);
    int n = 0;
    in >> n;
    if (n <= 0) throw std::runtime_error("FIRFilter: taps file has invalid format.");
    taps.resize(n);
    for (int i = 0; i < n; i++) {
        in >> taps[i];
        if (in.fail()) throw std::runtime_error("FIRFilter: file " + file + " should contain more taps");
    }
    return taps;
}

/* returns the dpd_mode and the coefficient block in dabmod_b200_config layout */
inline int load_coefs(std::istream& in, std::vector<float>& coefs)
{
    int fmt = 0;
    in >> fmt;
    if (fmt == 1) {
        int n = 0;
        in >> n;
        if (n != 5) throw std::runtime_error("MemlessPoly: invalid number of coefs: " + std::to_string(n));
        coefs.resize(10);
        for (auto& c : coefs) in >> c;
        if (in.fail()) throw std::runtime_error("MemlessPoly: coefs file invalid !");
        return DABMOD_B200_DPD_ODD_POLY;
    }
    if (fmt == 2) {
        coefs.resize(33);
        for (auto& c : coefs) in >> c;
        if (in.fail()) throw std::runtime_error("MemlessPoly: coefs file invalid !");
        return DABMOD_B200_DPD_LUT;
    }
    throw std::runtime_error("MemlessPoly: coef file has unknown format " + std::to_string(fmt));
}

inline int load_coefs(const std::string& file, std::vector<float>& coefs)
{
    std::ifstream in(file);
    if (!in) throw std::runtime_error("MemlessPoly: Could not open file with coefs!");
    return load_coefs(in, coefs);
}

inline int format_code(const std::string& format)
{
    if (format.empty() || format == "complexf") return DABMOD_B200_FMT_COMPLEXF;
    if (format == "s16") return DABMOD_B200_FMT_S16;
    if (format == "u8") return DABMOD_B200_FMT_U8;
    if (format == "s8") return DABMOD_B200_FMT_S8;
    throw std::runtime_error("FormatConverter: Invalid format " + format);
}

} // namespace b200files
