"""CPU restatement of the integer-residue (CRT) FP64 contraction of csrc/gemm_i8.cuh — test infrastructure.

C = A B^T for FP64 A (m x k), B (n x k), evaluated as the tcgen05 kind::i8 path does it ("Ozaki scheme II"):
  1. integerise: A'[i,:] = rint(A[i,:] * 2^(b - eA_i)), 2^eA_i > max_j |A_ij| (per row; likewise B), |A'| <= 2^b;
  2. residues: A_t = A' mod p_t in [0, p_t) (uint8) for T pairwise coprime moduli p_t <= 256;
  3. T independent u8 GEMMs with exact 32-bit accumulation, reduced mod p_t: R_t = (A_t B_t^T) mod p_t (uint8 again);
  4. CRT: C' = sum_t R_t w_t mod P  (w_t = (P/p_t) * ((P/p_t)^-1 mod p_t)) in 40-bit words whose partial sums are exact in
     FP64 (R_t <= 255, 16 terms, words < 2^40: sums < 2^52); the multiple of P is rint(total / P) evaluated in FP64;
  5. C = C' * 2^(eA_i + eB_j - 2b).
The integer product is EXACT; the only error is the truncation of the operands to b bits below their row maximum.
This file mirrors the device arithmetic step by step (same words, same order) with numpy int64 / float64, and offers the
same computation in exact Python integers for cross-checking."""
import numpy as np

MODULI = (256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193)
WORD_BITS = 40
N_WORDS = 4


def crt_constants(T):
    """Per modulus: the BYTES of the CRT weight w_t = (P / p_t) ((P / p_t)^-1 mod p_t) (16 bytes, little endian), the fraction
    w_t / P; and the 40-bit words of P."""
    ps = MODULI[:T]
    P = 1
    for p in ps:
        P *= p
    wbytes = np.zeros((T, 16), dtype=np.int64)
    frac = np.zeros(T, dtype=np.float64)
    for t, p in enumerate(ps):
        q = P // p
        w = q * pow(q % p, -1, p)
        frac[t] = w / P                                   # Python int true division: correctly rounded
        for k in range(16):
            wbytes[t, k] = (w >> (8 * k)) & 0xff
    Pw = np.array([float((P >> (WORD_BITS * k)) & ((1 << WORD_BITS) - 1)) for k in range(N_WORDS)])
    return P, wbytes, frac, Pw


def max_bits(T, k_red):
    """Largest b (bits per operand, both equal) such that 2 * k_red * 2^(2b) < P * (1 - 2^-30)."""
    P = crt_constants(T)[0]
    b = 0
    while 2 * k_red * (1 << (2 * (b + 1))) < P - (P >> 30):
        b += 1
    return min(b, 53)


def integerise(A, b):
    """rows of A -> (int64 A' with |A'| <= 2^b, exponents e): A ~= A' * 2^(e - b)."""
    amax = np.abs(A).max(axis=1)
    e = np.where(amax > 0, np.frexp(amax)[1], 0).astype(np.int64)        # 2^e > max |A_ij| (frexp: max = f 2^e, f in [0.5, 1))
    Ai = np.rint(np.ldexp(A, (b - e)[:, None])).astype(np.int64)
    return Ai, e


def residues(Ai, T):
    """(T, rows, cols) uint8 residues in [0, p) — what the device's byte-limb / multiply-high arithmetic yields."""
    out = np.empty((T,) + Ai.shape, dtype=np.uint8)
    for t, p in enumerate(MODULI[:T]):
        out[t] = np.mod(Ai, p).astype(np.uint8)          # numpy's mod is non-negative for a positive modulus
    return out


def gemm_mod(Ar, Br, T):
    """R_t = (A_t B_t^T) mod p_t in [0, p) (exact integer accumulation: u8 x u8 products, K * 255^2 < 2^31)."""
    out = np.empty((T, Ar.shape[1], Br.shape[1]), dtype=np.uint8)
    for t, p in enumerate(MODULI[:T]):
        acc = Ar[t].astype(np.int64) @ Br[t].astype(np.int64).T
        assert 0 <= acc.min() and acc.max() < 2 ** 31
        out[t] = np.mod(acc, p).astype(np.uint8)
    return out


def combine(R, eA, eB, bA, bB, T):
    """CRT reconstruction as the device does it (k_crt_combine) and scaling back: (rows, cols) float64.
    Integer byte sums S_k = sum_t r_t byte_k(w_t) (the device's dp4a), four 40-bit-spaced words W_j = sum_{i<5} S_{5j+i} 256^i
    (< 2^53: exact in FP64), m = rint(total / P) from the FP64 evaluation of the words, W_j - m P_j (exact), carries, and the
    final three multiply-adds (the only roundings)."""
    P, wbytes, frac, Pw = crt_constants(T)
    r = R.astype(np.int64)
    S = np.zeros((20,) + R.shape[1:], dtype=np.int64)
    for t in range(T):
        for k in range(16):
            S[k] += r[t] * wbytes[t, k]
    W = [sum(S[5 * j + i] << (8 * i) for i in range(5)).astype(np.float64) for j in range(N_WORDS)]
    two = float(1 << WORD_BITS)
    # the multiple of P: total / P in FP64 (|m| <= 2^12, error ~2^-41; the bit budget keeps |C'| / P away from 1/2).  The device
    # evaluates the chain with fused multiply-adds; every intermediate here is below 2^53 times a power of two only at the last
    # step, so the chain is replayed in exact integers and rounded once per fma
    m = np.rint(_fma_chain(W, two) * (1.0 / P))
    D = [W[k] - m * Pw[k] for k in range(N_WORDS)]     # exact: W < 2^53, m P_k < 2^52
    for k in range(N_WORDS - 1):                        # carry normalisation: |D_k| <= 2^39 for k < top
        c = np.rint(D[k] / two)
        D[k] = D[k] - c * two
        D[k + 1] = D[k + 1] + c
    return np.ldexp(_fma_chain(D, two), (eA[:, None] + eB[None, :] - bA - bB))


def _fma(a, b, c):
    """Element-wise fma(a, b, c) with ONE rounding, for integer-valued float64 arrays (exact Python integers in between)."""
    out = np.empty(a.shape, dtype=np.float64)
    af, cf, of = a.ravel(), c.ravel(), out.ravel()
    for i in range(af.size):
        of[i] = float(int(af[i]) * int(b) + int(cf[i]))          # int -> float: correctly rounded
    return out


def _fma_chain(words, two):
    val = words[N_WORDS - 1]
    for k in range(N_WORDS - 2, -1, -1):
        val = _fma(val, two, words[k])
    return val


def crt_matmul(A, B, T=16, bits=None):
    """A (m, k) @ B (n, k)^T through the residue pipeline."""
    b = max_bits(T, A.shape[1]) if bits is None else bits
    Ai, eA = integerise(A, b)
    Bi, eB = integerise(B, b)
    R = gemm_mod(residues(Ai, T), residues(Bi, T), T)
    return combine(R, eA, eB, b, b, T)


def exact_matmul_of_truncated(A, B, b):
    """Exact (Python integer) product of the b-bit truncated operands — what crt_matmul must equal up to final rounding."""
    Ai, eA = integerise(A, b)
    Bi, eB = integerise(B, b)
    C = Ai.astype(object) @ Bi.astype(object).T
    out = np.empty(C.shape)
    for i in range(C.shape[0]):
        for j in range(C.shape[1]):
            out[i, j] = float(C[i, j]) * 2.0 ** float(eA[i] + eB[j] - 2 * b) if abs(int(eA[i] + eB[j] - 2 * b)) < 1000 else 0.0
    return out
