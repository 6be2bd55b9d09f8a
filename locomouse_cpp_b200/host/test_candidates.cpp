// test_candidates.cpp — CPU checks of the Candidate / P22D value types against the behaviour of the
// reference's Candidates/Candidates.cpp (defaults, sentinel, overwrite-first, ordering).  Run by
// tests/test_host_cpp.py; prints "ok" and exits 0 when every check holds.
#include <algorithm>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "Candidates.hpp"

static int fails = 0;
#define CHECK(c)                                                    \
    do {                                                            \
        if (!(c)) {                                                 \
            std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
            ++fails;                                                \
        }                                                           \
    } while (0)

int main() {
    Candidate d;  // Candidates.cpp:4-7
    CHECK(d.p.x == -1 && d.p.y == -1 && d.s == -1);
    Candidate a(3, 4, 0.5), b(cv::Point_<int>(7, 8), 1.5);
    CHECK(a.point().x == 3 && a.point().y == 4 && a.score() == 0.5);
    CHECK(compareCandidate(b, a) && !compareCandidate(a, b) && !compareCandidate(a, a));  // score descending, strict
    std::vector<Candidate> v{a, b, Candidate(0, 0, 1.0)};
    std::sort(v.begin(), v.end(), compareCandidate);
    CHECK(v[0].s == 1.5 && v[1].s == 1.0 && v[2].s == 0.5);
    a.set_score(2.0);
    CHECK(a.score() == 2.0);

    P22D e;  // sentinel: Candidates.cpp:40-45, 148-156
    CHECK(e.number_of_candidates() == 0 && e.y_side_coord(0) == -1 && e.score_side(0) == -1);
    CHECK(e.x_coord() == -1 && e.y_bottom_coord() == -1 && e.score_bottom() == -1);

    P22D none(Candidate(10, 20, 0.9), Candidate(-1, -1, -1));  // matchViews' "no side match" record
    CHECK(none.number_of_candidates() == 0 && none.x_coord() == 10 && none.y_bottom_coord() == 20 && none.score_bottom() == 0.9);
    none.add_side_candidate(33, 0.25);  // first real candidate overwrites the sentinel slot (Candidates.cpp:106-111)
    CHECK(none.number_of_candidates() == 1 && none.y_side_coord(0) == 33 && none.score_side(0) == 0.25);
    none.add_side_candidate(Candidate(99, 44, 0.0));  // zero score is valid; x of a side candidate is ignored
    CHECK(none.number_of_candidates() == 2 && none.y_side_coord(1) == 44 && none.point_side(1).x == 10);
    CHECK(none.get_candidate_side(1).p.x == 10 && none.get_candidate_side(1).p.y == 44 && none.get_candidate_side(1).s == 0.0);
    bool threw = false;
    try {
        none.add_side_candidate(5, -0.5);  // CV_Assert(S >= 0) in the reference
    } catch (const std::runtime_error &) {
        threw = true;
    }
    CHECK(threw && none.number_of_candidates() == 2);

    P22D two(1, 2, 3, 0.5, 0.75);
    CHECK(two.number_of_candidates() == 1 && two.point_bottom() == cv::Point_<int>(1, 2) && two.point_side(0) == cv::Point_<int>(1, 3));
    P22D pp(cv::Point_<int>(5, 6), cv::Point_<int>(50, 7), 0.1, 0.2);
    CHECK(pp.x_coord() == 5 && pp.y_side_coord(0) == 7 && pp.score_side(0) == 0.2);
    // a negative first side score keeps the record "empty" (st[0] < 0), and the next add overwrites it
    P22D neg(Candidate(1, 1, 1.0), Candidate(1, 9, -0.1));
    CHECK(neg.number_of_candidates() == 0);
    neg.add_side_candidate(8, 0.3);
    CHECK(neg.number_of_candidates() == 1 && neg.y_side_coord(0) == 8);

    std::ostringstream os;
    os << none << a;
    two.write(os);
    CHECK(os.str().find("2 top candidate(s)") != std::string::npos && os.str().find("n_candidates_side: 1") != std::string::npos);
    if (!fails) std::printf("ok\n");
    return fails ? 1 : 0;
}
