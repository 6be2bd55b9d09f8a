// LocoMouse_ParseInputs — the reference's 8 positional command-line inputs, same member names and
// order as LocoMouse_Core/LocoMouse_ParseInputs.hpp:19-28 / LocoMouse_ParseInputs.cpp:56-100:
//   LocoMouse method config.yml video background model calibration side_char output_folder
#pragma once
#include <string>

class LocoMouse_ParseInputs {
public:
    std::string LM_CALL, CONFIG_FILE, VIDEO_FILE, MODEL_FILE, BKG_FILE, CALIBRATION_FILE, FLIP_CHAR, OUTPUT_PATH, METHOD,
        REF_PATH, FILE_STEM;

    LocoMouse_ParseInputs() = default;
    LocoMouse_ParseInputs(int argc, char *argv[]);  // throws std::invalid_argument on a wrong count
    std::string stripFileName(const std::string &s);
};
