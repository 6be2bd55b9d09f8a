// Candidates.hpp — the value types the detection path produces and the host tracker consumes.
// Same public interface as the reference's Candidates/Candidates.hpp:16-105 (Candidate, P22D,
// compareCandidate, stream output); behaviour follows Candidates/Candidates.cpp:4-156, in particular
// the "no side match" sentinel of P22D.  OpenCV FileStorage (de)serialisers are replaced
// by a plain YAML-text writer (file formats are outside the hot path).
#pragma once
#include <iosfwd>
#include <vector>

#include "cv_shim.hpp"

typedef unsigned int uint;

class Candidate {
public:
    cv::Point_<int> p;
    double s;

    Candidate();                       // (-1,-1), score -1   (Candidates.cpp:4-7)
    Candidate(int x, int y, double score);
    Candidate(cv::Point_<int> point, double score);

    cv::Point_<int> point() const { return p; }
    double score() const { return s; }
    void set_score(double new_s) { s = new_s; }

    void write(std::ostream &os) const;  // {Point_x: .., Point_y: .., Score: ..}
};

// score-descending order used by the reference's std::sort calls (Candidates.cpp:33-36)
bool compareCandidate(Candidate a, Candidate b);
std::ostream &operator<<(std::ostream &out, const Candidate &c);

// One bottom-view candidate with zero or more side-view (y, score) matches.
class P22D {
    Candidate CB;
    std::vector<int> yt;     // side-view rows; yt[0] == -1 with st[0] == -1 means "none"
    std::vector<double> st;  // side-view scores

public:
    P22D();
    P22D(int x, int y_bottom, int y_side, double score_bottom, double score_side);
    P22D(cv::Point_<int> p_bottom, cv::Point_<int> p_side, double score_bottom, double score_side);
    P22D(Candidate c_bottom, Candidate c_side);

    cv::Point_<int> point_bottom() const;
    cv::Point_<int> point_side(uint i) const;
    double score_bottom() const;
    double score_side(uint i) const;
    int x_coord() const;
    int y_bottom_coord() const;
    int y_side_coord(uint i) const;

    void add_side_candidate(Candidate c);
    void add_side_candidate(cv::Point_<int> p, double s);
    void add_side_candidate(int y, double s);

    int number_of_candidates() const;  // 0 when only the sentinel is stored
    Candidate get_candidate_side(uint i) const;
    Candidate get_candidate_bottom() const;

    void write(std::ostream &os) const;

private:
    void add_side_candidate_safe(int x, int y, double s);
};
std::ostream &operator<<(std::ostream &out, const P22D &c);
