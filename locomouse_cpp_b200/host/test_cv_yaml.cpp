// test_cv_yaml.cpp — dumps what cv_yaml.hpp reads from an OpenCV YAML file, for tests/test_host_cpp.py (the file is written
// by the real OpenCV through cv2.FileStorage).  usage: test_cv_yaml file key...   prints  key kind rows cols dt values...
#include <cstdio>

#include "cv_yaml.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    try {
        cvyaml::File y(argv[1]);
        if (!y.isOpened()) {
            std::printf("not-opened\n");
            return 1;
        }
        for (int i = 2; i < argc; ++i) {
            const cvyaml::Matrix &m = y.mat(argv[i]);
            if (!m.empty()) {
                std::printf("%s matrix %d %d %c", argv[i], m.rows, m.cols, m.dt);
                for (double v : m.data) std::printf(" %.17g", v);
                std::printf("\n");
            } else if (y.has(argv[i])) {
                double v = 0.0;
                bool numeric = true;
                try {
                    v = y.real(argv[i]);
                } catch (const std::exception &) {
                    numeric = false;
                }
                if (numeric) std::printf("%s scalar %.17g [%s]\n", argv[i], v, y.str(argv[i]).c_str());
                else std::printf("%s string - [%s]\n", argv[i], y.str(argv[i]).c_str());
            } else {
                std::printf("%s missing %.17g\n", argv[i], y.real(argv[i]));
            }
        }
    } catch (const std::exception &e) {
        std::printf("error %s\n", e.what());
        return 3;
    }
    return 0;
}
