// Load / store modes of the line engine (shared by host and device code).
#pragma once
namespace rfb {
// how a line is read
enum : int {
    LD_C2C = 0,   // complex
    LD_REAL = 1,  // real, imaginary part 0
    LD_HERM = 2,  // complex half spectrum (bins 0..n/2) expanded by Hermitian symmetry
    LD_HC = 3,    // FFTPACK halfcomplex real line expanded to the full spectrum
    LD_DCT2 = 4,  // fused DCT-II / DST-II  (power-of-two kernel only; pairs with ST_DCT2)
    LD_DCT3 = 5,  // fused DCT-III / DST-III (power-of-two kernel only; pairs with ST_DCT3)
};
// how a line is written
enum : int {
    ST_C2C = 0,      // complex, all bins
    ST_HALF = 1,     // complex, bins 0..n/2
    ST_REAL = 2,     // real part
    ST_HC = 3,       // FFTPACK halfcomplex packing of bins 0..n/2
    ST_HARTLEY = 4,  // Re + Im
    ST_DCT2 = 5,
    ST_DCT3 = 6,
};
enum : int {
    FLAG_NEG_EVEN_IN = 1,
    FLAG_NEG_EVEN_OUT = 2,
    FLAG_SINE = 4,    // DST instead of DCT
    FLAG_ORTHO = 8,   // orthogonalised variant
    FLAG_QUIRK = 16,  // DST-II/III ortho scaling on element 0 like the reference (H:3033-3039)
};
}  // namespace rfb
