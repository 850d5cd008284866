// term.h -- one explicit (RHS) term: a vector of prefactor monomials times a product of fields.
// Mirrors the public face of /root/reference/inc/cupss/term.h:38-60.  In this implementation a term is
// pure description: the engine fuses all terms of a sweep into its kernels, so there are no per-term
// device arrays, FFT plans or prefactor tables.
#ifndef CUPSS_B200_TERM_H
#define CUPSS_B200_TERM_H

#include <map>
#include <string>
#include <vector>
#include "defines.h"

class field;

class term {
   public:
    term(int sx, float dx);
    term(int sx, int sy, float dx, float dy);
    term(int sx, int sy, int sz, float dx, float dy, float dz);
    ~term();

    bool isCUDA = true;
    pres prefactors;                        // unused legacy slot kept for source compatibility
    std::vector<pres> prefactors_h;         // the monomials
    std::vector<field *> product;           // factors (empty: the constant 1)
    int multiply_by_i_pre = 0;              // odd total power of i
    dim3 threads_per_block, blocks;

    // runtime parameter updates
    std::map<std::string, int> usedParameters;
    std::vector<std::string> prefactor_strings;
    int setPrefactorString(const std::vector<std::string> &strings);
    void printPrefactorString();

    int prepareDevice();          // computes multiply_by_i_pre and flags the product fields for dealiasing
    int precomputePrefactors();   // nothing to tabulate; validates the powers of i
    int update();                 // folded into evolver::advanceTime; kept as a no-op

   private:
    const int sx, sy, sz;
    const float dx, dy, dz;
};

#endif
