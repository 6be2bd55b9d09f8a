// MyMat.hpp — value types with the interface of the reference's MyMat / MATSPARSE (MyMat/MyMat.hpp:16-91): a dense
// column-major double matrix and a MATLAB-style CSC sparse matrix, as the host tracker (match2nd) consumes them.  Here they
// are filled from the device-built cost arrays (lm_unary_costs / lm_pairwise_costs) instead of element by element.
#pragma once
#include <cassert>
#include <vector>

class MyMat {
    std::vector<double> values;
    int nrows = 0, ncols = 0;

public:
    MyMat() = default;
    MyMat(unsigned int n, unsigned int m) : values((size_t)n * m, 0.0), nrows((int)n), ncols((int)m) {}
    void put(unsigned int i, unsigned int j, double v) {
        assert((int)i < nrows && (int)j < ncols);
        values[(size_t)j * nrows + i] = v;
    }
    double get(unsigned int i, unsigned int j) const {
        assert((int)i < nrows && (int)j < ncols);
        return values[(size_t)j * nrows + i];
    }
    int Nrows() const { return nrows; }
    int Ncols() const { return ncols; }
    int Numel() const { return nrows * ncols; }
    const double *getValues() const { return values.data(); }
    double *getValues() { return values.data(); }
};

class MATSPARSE {
    std::vector<int> Ir, Jc;
    std::vector<double> Pr;
    int n_rows = 0, n_cols = 0;

public:
    MATSPARSE() = default;
    // MATSPARSE(const MyMat*) (MyMat.cpp:141-178): column by column, entries equal to 0 are not stored
    explicit MATSPARSE(const MyMat *M) : n_rows(M->Nrows()), n_cols(M->Ncols()) {
        Jc.push_back(0);
        for (int j = 0; j < n_cols; ++j) {
            for (int i = 0; i < n_rows; ++i) {
                const double v = M->get((unsigned int)i, (unsigned int)j);
                if (v != 0) {
                    Ir.push_back(i);
                    Pr.push_back(v);
                }
            }
            Jc.push_back((int)Ir.size());
        }
    }
    MATSPARSE(int rows, int cols, const int *jc, const int *ir, const double *pr)
        : Ir(ir, ir + jc[cols]), Jc(jc, jc + cols + 1), Pr(pr, pr + jc[cols]), n_rows(rows), n_cols(cols) {}
    const int *getIr() const { return Ir.data(); }
    const int *getJc() const { return Jc.data(); }
    const double *getPr() const { return Pr.data(); }
    int nz() const { return (int)Ir.size(); }
    int Nrows() const { return n_rows; }
    int Ncols() const { return n_cols; }
};
