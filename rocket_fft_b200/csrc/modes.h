// Load / store modes of the line engine (shared by host and device code).
#pragma once
namespace rfb {
// how a line is read
enum : int {
    LD_C2C = 0,   // complex
    LD_REAL = 1,  // real, imaginary part 0
    LD_HERM = 2,  // complex half spectrum (bins 0..n/2) expanded by Hermitian symmetry
    LD_HC = 3,    // FFTPACK halfcomplex real line expanded to the full spectrum
};
// how a line is written
enum : int {
    ST_C2C = 0,      // complex, all bins
    ST_HALF = 1,     // complex, bins 0..n/2
    ST_REAL = 2,     // real part
    ST_HC = 3,       // FFTPACK halfcomplex packing of bins 0..n/2
    ST_HARTLEY = 4,  // Re + Im
};
enum : int { FLAG_NEG_EVEN_IN = 1, FLAG_NEG_EVEN_OUT = 2 };
}  // namespace rfb
