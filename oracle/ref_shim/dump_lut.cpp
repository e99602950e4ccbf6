// TEST INFRASTRUCTURE ONLY (oracle/): dumps the reference's shipped decode tables
// (src/dotp_lut.hpp: dotp_lut_a, dotp_lut_b; src/na_lut.hpp: na_lut) to a flat binary so
// tests/golden/ can hold them as the bit-exact decode known-answer vectors.
#include <cstdio>
#include "dotp_lut.hpp"
#include "na_lut.hpp"
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "wb");
    if (!f) return 1;
    fwrite(dotp_lut_a, sizeof(double), 1024, f);
    fwrite(dotp_lut_b, sizeof(double), 1024, f);
    fwrite(na_lut, sizeof(double), 64, f);
    fclose(f);
    return 0;
}
