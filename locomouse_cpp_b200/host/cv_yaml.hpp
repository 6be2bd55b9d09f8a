// cv_yaml.hpp — reader for the subset of OpenCV's YAML 1.0 persistence format that the reference's input files use
// (cv::FileStorage READ in LocoMouse_class.cpp:12-200 config, 419-463 calibration, 3095-3162 model): top-level
//     key: scalar
//     key: !!opencv-matrix
//        rows: R
//        cols: C
//        dt: d | f | i | u | s | w | c          (single channel)
//        data: [ v, v, ... ]                     (flow sequence, may span many lines)
// The OpenCV C++ SDK is not available in this image, so FileStorage itself cannot be used; files written by the real
// OpenCV (cv2.FileStorage) are what tests/test_host_cpp.py feeds this reader.  Host-side I/O only.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cvyaml {

struct Matrix {
    int rows = 0, cols = 0;
    char dt = 'd';
    std::vector<double> data;  // row-major; every supported depth is exactly representable in a double
    bool empty() const { return rows <= 0 || cols <= 0; }
    bool is_integer() const { return dt == 'i' || dt == 'u' || dt == 's' || dt == 'w' || dt == 'c'; }
};

class File {
    std::map<std::string, std::string> scalars_;
    std::map<std::string, Matrix> mats_;
    bool opened_ = false;

    static std::string trim(const std::string &s) {
        const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    static double number(const std::string &t, const std::string &where) {
        if (t == ".Inf" || t == ".inf" || t == "+.Inf") return std::numeric_limits<double>::infinity();
        if (t == "-.Inf" || t == "-.inf") return -std::numeric_limits<double>::infinity();
        if (t == ".Nan" || t == ".NaN" || t == ".nan") return std::numeric_limits<double>::quiet_NaN();
        char *end = nullptr;
        const double v = std::strtod(t.c_str(), &end);
        if (end == t.c_str() || (*end != '\0' && *end != '.')) throw std::invalid_argument("Cannot parse the number '" + t + "' in " + where);
        return v;
    }

public:
    explicit File(const std::string &name) {
        std::ifstream in(name);
        if (!in) return;
        std::string first;
        std::getline(in, first);
        if (first.rfind("%YAML", 0) != 0) return;  // not an OpenCV YAML file
        opened_ = true;
        std::string line, key;
        Matrix *cur = nullptr;
        bool in_data = false;
        std::string data;
        auto finish_data = [&]() {
            std::string tok;
            for (char ch : data) {
                if (ch == ',' || ch == '[' || ch == ']' || ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') {
                    if (!tok.empty()) cur->data.push_back(number(tok, name + " (" + key + ")"));
                    tok.clear();
                } else
                    tok += ch;
            }
            if (!tok.empty()) cur->data.push_back(number(tok, name + " (" + key + ")"));
            if ((long long)cur->data.size() != (long long)cur->rows * cur->cols)
                throw std::invalid_argument("Matrix " + key + " in " + name + " holds " + std::to_string(cur->data.size()) + " values, " +
                                            std::to_string((long long)cur->rows * cur->cols) + " expected (multi-channel data is not supported).");
            data.clear();
            in_data = false;
            cur = nullptr;
        };
        while (std::getline(in, line)) {
            if (in_data) {
                data += line;
                data += '\n';
                if (line.find(']') != std::string::npos) finish_data();
                continue;
            }
            if (line.rfind("---", 0) == 0 || line.rfind("...", 0) == 0 || trim(line).empty() || trim(line)[0] == '#') continue;
            const size_t colon = line.find(':');
            if (colon == std::string::npos) continue;
            const bool indented = line[0] == ' ' || line[0] == '\t';
            const std::string k = trim(line.substr(0, colon)), v = trim(line.substr(colon + 1));
            if (!indented) {
                key = k;
                cur = nullptr;
                if (v.rfind("!!opencv-matrix", 0) == 0) {
                    cur = &mats_[key];
                    *cur = Matrix();
                } else {
                    std::string s = v;
                    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) s = s.substr(1, s.size() - 2);
                    scalars_[key] = s;
                }
            } else if (cur) {
                if (k == "rows") cur->rows = std::atoi(v.c_str());
                else if (k == "cols") cur->cols = std::atoi(v.c_str());
                else if (k == "dt") {
                    if (v.size() != 1) throw std::invalid_argument("Matrix " + key + " in " + name + ": element type '" + v + "' is not supported.");
                    cur->dt = v[0];
                } else if (k == "data") {
                    cur->data.reserve((size_t)std::max(0, cur->rows) * std::max(0, cur->cols));
                    in_data = true;
                    data = v;
                    data += '\n';
                    if (v.find(']') != std::string::npos) finish_data();
                }
            }
        }
        if (in_data) throw std::invalid_argument("Unterminated data sequence of " + key + " in " + name);
    }
    bool isOpened() const { return opened_; }
    bool has(const std::string &key) const { return scalars_.count(key) || mats_.count(key); }
    // FileNode >> double semantics: a missing node leaves 0 (the reference relies on this for the biases, class.cpp:3140)
    double real(const std::string &key) const {
        auto it = scalars_.find(key);
        return it == scalars_.end() ? 0.0 : number(it->second, key);
    }
    std::string str(const std::string &key) const {
        auto it = scalars_.find(key);
        return it == scalars_.end() ? std::string() : it->second;
    }
    const Matrix &mat(const std::string &key) const {  // a missing node reads as an empty matrix, as FileNode >> Mat does
        static const Matrix none;
        auto it = mats_.find(key);
        return it == mats_.end() ? none : it->second;
    }
};


// Writer for the reference's output file (cv::FileStorage WRITE, LocoMouse_class.cpp:360, 2385-2482): top-level
// 32-bit integer matrices as `name: !!opencv-matrix` nodes.  The flow sequence is wrapped exactly as OpenCV's YAML
// emitter wraps it (an item moves to a new line, indented by 7, when the line would pass column 71), so files are
// byte-identical to what cv2.FileStorage writes for the same matrices (tests/test_host_tracks.py).
class Writer {
    std::FILE *f_ = nullptr;  // C stdio: this header is also built into a dlopen()ed test library
    bool ok_ = true;
    void put(const std::string &t) { ok_ = ok_ && std::fwrite(t.data(), 1, t.size(), f_) == t.size(); }

public:
    explicit Writer(const std::string &name) : f_(std::fopen(name.c_str(), "wb")) {
        if (!f_) throw std::runtime_error("Could not open the output file: " + name);
        put("%YAML:1.0\n---\n");
    }
    Writer(const Writer &) = delete;
    Writer &operator=(const Writer &) = delete;
    ~Writer() {
        if (f_) std::fclose(f_);
    }
    void write(const std::string &name, int rows, int cols, const int *data) {
        put(name + ": !!opencv-matrix\n   rows: " + std::to_string(rows) + "\n   cols: " + std::to_string(cols) + "\n   dt: i\n");
        std::string line = "   data: [";
        bool first = true;
        for (long i = 0; i < (long)rows * cols; ++i) {
            const std::string item = std::to_string(data[i]);
            if (!first) line += ',';
            if (line.size() + item.size() > 71 && line.size() > 7) {
                put(line + "\n");
                line.assign(7, ' ');
            } else {
                line += ' ';
            }
            line += item;
            first = false;
        }
        put(line + " ]\n");
    }
    bool good() {
        ok_ = ok_ && std::fflush(f_) == 0;
        return ok_;
    }
};

}  // namespace cvyaml
