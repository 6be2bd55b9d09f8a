// Candidates.cpp — see Candidates.hpp.  Semantics restated from Candidates/Candidates.cpp:4-156.
#include "Candidates.hpp"

#include <cassert>
#include <ostream>
#include <stdexcept>

Candidate::Candidate() : p(-1, -1), s(-1) {}
Candidate::Candidate(int x, int y, double score) : p(x, y), s(score) {}
Candidate::Candidate(cv::Point_<int> point, double score) : p(point), s(score) {}

void Candidate::write(std::ostream &os) const {
    os << "{ Point_x: " << p.x << ", Point_y: " << p.y << ", Score: " << s << " }";
}

bool compareCandidate(Candidate a, Candidate b) { return a.score() > b.score(); }

std::ostream &operator<<(std::ostream &out, const Candidate &c) {
    return out << "[(" << c.p.x << ", " << c.p.y << ") with score = " << c.s << "]";
}

// A default P22D carries the sentinel side entry (-1, -1): Candidates.cpp:40-45.
P22D::P22D() : CB(), yt(1, -1), st(1, -1.0) {}
P22D::P22D(int x, int y_bottom, int y_side, double score_bottom, double score_side)
    : CB(x, y_bottom, score_bottom), yt(1, y_side), st(1, score_side) {}
P22D::P22D(cv::Point_<int> p_bottom, cv::Point_<int> p_side, double score_bottom, double score_side)
    : CB(p_bottom, score_bottom), yt(1, p_side.y), st(1, score_side) {}
P22D::P22D(Candidate c_bottom, Candidate c_side)
    : CB(c_bottom.point(), c_bottom.score()), yt(1, c_side.point().y), st(1, c_side.score()) {}

cv::Point_<int> P22D::point_bottom() const { return CB.point(); }
double P22D::score_bottom() const { return CB.score(); }
int P22D::x_coord() const { return CB.point().x; }
int P22D::y_bottom_coord() const { return CB.point().y; }

void P22D::add_side_candidate(Candidate c) { add_side_candidate_safe(c.point().x, c.point().y, c.score()); }
void P22D::add_side_candidate(cv::Point_<int> p, double s) { add_side_candidate_safe(p.x, p.y, s); }
void P22D::add_side_candidate(int y, double s) { add_side_candidate_safe(CB.point().x, y, s); }

// The first real side candidate overwrites the sentinel slot; later ones must have a non-negative
// score (the reference CV_Asserts, Candidates.cpp:106-115 — an uncaught cv::Exception there; a
// std::runtime_error here so main()'s catch reports it).
void P22D::add_side_candidate_safe(int, int y, double s) {
    if (number_of_candidates() == 0) {
        yt[0] = y;
        st[0] = s;
        return;
    }
    if (!(s >= 0)) throw std::runtime_error("P22D::add_side_candidate: negative side score");
    yt.push_back(y);
    st.push_back(s);
}

cv::Point_<int> P22D::point_side(uint i) const {
    assert(i < yt.size());
    return cv::Point_<int>(CB.point().x, yt[i]);
}
int P22D::y_side_coord(uint i) const {
    assert(i < yt.size());
    return yt[i];
}
double P22D::score_side(uint i) const {
    assert(i < st.size());
    return st[i];
}
Candidate P22D::get_candidate_side(uint i) const {
    assert(i < yt.size());
    return Candidate(CB.point().x, yt[i], st[i]);
}
Candidate P22D::get_candidate_bottom() const { return CB; }

int P22D::number_of_candidates() const { return st[0] < 0 ? 0 : (int)st.size(); }

void P22D::write(std::ostream &os) const {
    os << "{ Candidate_bottom: ";
    CB.write(os);
    os << ", n_candidates_side: " << yt.size() << ", Candidates_side: [";
    for (size_t i = 0; i < yt.size(); ++i) os << (i ? ", " : "") << yt[i];
    os << "], Scores_side: [";
    for (size_t i = 0; i < st.size(); ++i) os << (i ? ", " : "") << st[i];
    os << "] }";
}

std::ostream &operator<<(std::ostream &out, const P22D &c) {
    out << "Bottom candidate: " << c.get_candidate_bottom() << "\n" << c.number_of_candidates() << " top candidate(s):\n";
    for (int i = 0; i < c.number_of_candidates(); ++i)
        out << "[" << c.y_side_coord(i) << " with score = " << c.score_side(i) << "]\n";
    return out;
}
