// term.cpp -- explicit-term descriptor (see inc/cupss/term.h).
#include "../../inc/cupss.h"

term::term(int nx, float hx) : sx(nx), sy(1), sz(1), dx(hx), dy(1.0f), dz(1.0f) {}
term::term(int nx, int ny, float hx, float hy) : sx(nx), sy(ny), sz(1), dx(hx), dy(hy), dz(1.0f) {}
term::term(int nx, int ny, int nz, float hx, float hy, float hz) : sx(nx), sy(ny), sz(nz), dx(hx), dy(hy), dz(hz) {}
term::~term() {}

int term::setPrefactorString(const std::vector<std::string> &strings) {
    prefactor_strings.insert(prefactor_strings.end(), strings.begin(), strings.end());
    return 0;
}

void term::printPrefactorString() {
    for (const std::string &s : prefactor_strings) std::cout << s << std::endl;
    for (const auto &kv : usedParameters) std::cout << kv.first << " " << kv.second << std::endl;
}

// Products of two or more fields (and the empty product) are formed in real space from the dealiased
// copies of their factors: flag those fields (term::prepareDevice, /root/reference/src/term_init.cpp:118-126).
int term::prepareDevice() {
    if (product.size() != 1)
        for (field *f : product) {
            f->needsaliasing = true;
            if (f->aliasing_order < (int)product.size()) f->aliasing_order = (int)product.size();
        }
    return precomputePrefactors();
}

int term::precomputePrefactors() {
    if (prefactors_h.empty()) return 0;
    const int parity = (prefactors_h[0].iqx + prefactors_h[0].iqy + prefactors_h[0].iqz) % 2;
    for (const pres &p : prefactors_h)
        if ((p.iqx + p.iqy + p.iqz) % 2 != parity) std::cout << "PANIC: Inconsistent powers of I in prefactors" << std::endl;
    multiply_by_i_pre = parity;
    return 0;
}

int term::update() { return 0; }
