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
    // DCT / DST of any length as the load / store stage of ONE complex transform (generic tile kernel, line_io.cuh); N = real
    // line length, n = transform length; each pairs with the ST_G_* of the same name
    LD_G_DCT1 = 6,   // n = 2(N-1): even extension of x                                  (T_dct1, H:2918-2955)
    LD_G_DST1 = 7,   // n = 2(N+1): odd extension of x                                   (T_dst1, H:2957-2985)
    LD_G_DCT2 = 8,   // n = N: Makhoul's reordering [x0, x2, .. , x3, x1]                  (T_dcst23, H:2987-3061)
    LD_G_DCT3 = 9,   // n = N, backward: (x_k - i x_{N-k}) e^{+i pi k / 2N}
    LD_G_DCT4 = 10,  // n = N/2 (N even): (x_{2j} + i x_{N-1-2j}) e^{-i pi (4j+1) / 4N}   (T_dcst4, H:3063-3163)
    LD_G_DCT4Z = 11, // n = 2N (any N): x_j e^{-i pi j / 2N}, zero padded
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
    ST_G_DCT1 = 7,   // y_k = Re X_k, k < N
    ST_G_DST1 = 8,   // y_k = -Im X_{k+1}, k < N
    ST_G_DCT2 = 9,   // y_k = 2 Re(e^{-i pi k / 2N} V_k)
    ST_G_DCT3 = 10,  // y_{2n} = v_n, y_{2n+1} = v_{N-1-n}
    ST_G_DCT4 = 11,  // W_k = Z_k e^{-i pi k / N}: y_{2k} = 2 Re W_k, y_{N-1-2k} = -2 Im W_k
    ST_G_DCT4Z = 12, // y_k = 2 Re(e^{-i pi (2k+1) / 4N} Z_k), k < N
};
enum : int {
    FLAG_NEG_EVEN_IN = 1,
    FLAG_NEG_EVEN_OUT = 2,
    FLAG_SINE = 4,    // DST instead of DCT
    FLAG_ORTHO = 8,   // orthogonalised variant
    FLAG_QUIRK = 16,  // DST-II/III ortho scaling on element 0 like the reference (H:3033-3039)
};
}  // namespace rfb
