"""Independent pure-Python restatements of the two algorithms whose long paths no reference vector pins
(CityHash128 v1.1 and the reference's rapidhash port).  Written separately from oracle/*.c and the CUDA code,
from the published algorithms, so a transcription slip in one of the three shows up as a disagreement.
Small inputs only (pure Python)."""
M = (1 << 64) - 1
K0, K1, K2, KMUL = 0xc3a5c85c97cb3127, 0xb492b66fbe98f273, 0x9ae16a3b2f90404f, 0x9ddfea08eb382d69


def _f64(b, i): return int.from_bytes(b[i:i + 8], "little")
def _f32(b, i): return int.from_bytes(b[i:i + 4], "little")
def _ror(v, s): return v if s == 0 else ((v >> s) | (v << (64 - s))) & M
def _smix(v): return v ^ (v >> 47)


def _h16(u, v, mul=KMUL):
    a = ((u ^ v) * mul) & M
    a ^= a >> 47
    b = ((v ^ a) * mul) & M
    b ^= b >> 47
    return (b * mul) & M


def _len0to16(s):
    n = len(s)
    if n >= 8:
        mul = (K2 + n * 2) & M
        a = (_f64(s, 0) + K2) & M
        b = _f64(s, n - 8)
        c = (_ror(b, 37) * mul + a) & M
        d = ((_ror(a, 25) + b) * mul) & M
        return _h16(c, d, mul)
    if n >= 4:
        mul = (K2 + n * 2) & M
        return _h16((n + (_f32(s, 0) << 3)) & M, _f32(s, n - 4), mul)
    if n > 0:
        y = (s[0] + (s[n >> 1] << 8)) & 0xFFFFFFFF
        z = (n + (s[n - 1] << 2)) & 0xFFFFFFFF
        return (_smix((y * K2 ^ z * K0) & M) * K2) & M
    return K2


def _weak(w, x, y, z, a, b):
    a = (a + w) & M
    b = _ror((b + a + z) & M, 21)
    c = a
    a = (a + x + y) & M
    b = (b + _ror(a, 44)) & M
    return (a + z) & M, (b + c) & M


def _weak_at(s, i, a, b): return _weak(_f64(s, i), _f64(s, i + 8), _f64(s, i + 16), _f64(s, i + 24), a, b)


def _city_murmur(s, lo, hi):
    n = len(s)
    a, b = lo, hi
    if n <= 16:
        a = (_smix((a * K1) & M) * K1) & M
        c = (b * K1 + _len0to16(s)) & M
        d = _smix((a + (_f64(s, 0) if n >= 8 else c)) & M)
    else:
        c = _h16((_f64(s, n - 8) + K1) & M, a)
        d = _h16((b + n) & M, (c + _f64(s, n - 16)) & M)
        a = (a + d) & M
        i, l = 0, n - 16
        while True:
            a ^= (_smix((_f64(s, i) * K1) & M) * K1) & M
            a = (a * K1) & M
            b ^= a
            c ^= (_smix((_f64(s, i + 8) * K1) & M) * K1) & M
            c = (c * K1) & M
            d ^= c
            i += 16
            l -= 16
            if l <= 0:
                break
    a = _h16(a, c)
    b = _h16(d, b)
    return a ^ b, _h16(b, a)


def _with_seed(s, lo, hi):
    n = len(s)
    if n < 128:
        return _city_murmur(s, lo, hi)
    x, y, z = lo, hi, (n * K1) & M
    v0 = (_ror(y ^ K1, 49) * K1 + _f64(s, 0)) & M
    v1 = (_ror(v0, 42) * K1 + _f64(s, 8)) & M
    w0 = (_ror((y + z) & M, 35) * K1 + x) & M
    w1 = (_ror((x + _f64(s, 88)) & M, 53) * K1) & M
    i = 0
    while True:
        for _ in range(2):
            x = (_ror((x + y + v0 + _f64(s, i + 8)) & M, 37) * K1) & M
            y = (_ror((y + v1 + _f64(s, i + 48)) & M, 42) * K1) & M
            x ^= w1
            y = (y + v0 + _f64(s, i + 40)) & M
            z = (_ror((z + w0) & M, 33) * K1) & M
            v0, v1 = _weak_at(s, i, (v1 * K1) & M, (x + w0) & M)
            w0, w1 = _weak_at(s, i + 32, (z + w1) & M, (y + _f64(s, i + 16)) & M)
            z, x = x, z
            i += 64
        n -= 128
        if n < 128:
            break
    x = (x + _ror((v0 + z) & M, 49) * K0) & M
    y = (y * K0 + _ror(w1, 37)) & M
    z = (z * K0 + _ror(w0, 27)) & M
    w0 = (w0 * 9) & M
    v0 = (v0 * K0) & M
    done = 0
    while done < n:
        done += 32
        y = (_ror((x + y) & M, 42) * K0 + v1) & M
        w0 = (w0 + _f64(s, i + n - done + 16)) & M
        x = (x * K0 + w0) & M
        z = (z + w1 + _f64(s, i + n - done)) & M
        w1 = (w1 + v0) & M
        v0, v1 = _weak_at(s, i + n - done, (v0 + z) & M, v1)
        v0 = (v0 * K0) & M
    x = _h16(x, v0)
    y = _h16((y + z) & M, w0)
    return (_h16((x + v1) & M, w1) + y) & M, _h16((x + w1) & M, (y + v1) & M)


def cityhash128_hex(s: bytes) -> str:
    """seqhasher's print order: High then Low (seqhasher.go:316)."""
    if len(s) >= 16:
        lo, hi = _with_seed(s[16:], _f64(s, 0), (_f64(s, 8) + K0) & M)
    else:
        lo, hi = _with_seed(s, K0, K1)
    return f"{hi:016x}{lo:016x}"


RS = (0x2d358dccaa6c78a5, 0x8bb84b93962eacc9, 0x4b33a62ed433d4a3)


def _mum(a, b):
    p = a * b
    return p & M, p >> 64


def _mix(a, b):
    lo, hi = _mum(a, b)
    return lo ^ hi


def rapidhash_hex(key: bytes, final_secret: int = 2) -> str:
    n = len(key)
    seed = 0xbdd89aa982704029
    seed ^= _mix(seed ^ RS[0], RS[1]) ^ n
    if n <= 16:
        if n >= 4:
            a = (_f32(key, 0) << 32) | _f32(key, n - 4)
            delta = (n & 24) >> (n >> 3)
            b = (_f32(key, delta) << 32) | _f32(key, n - 4 - delta)
        elif n > 0:
            a, b = (key[0] << 56) | (key[n >> 1] << 32) | key[n - 1], 0
        else:
            a = b = 0
    else:
        i, p = n, 0
        if i > 48:
            s1 = s2 = seed
            while True:
                seed = _mix(_f64(key, p) ^ RS[0], _f64(key, p + 8) ^ seed)
                s1 = _mix(_f64(key, p + 16) ^ RS[1], _f64(key, p + 24) ^ s1)
                s2 = _mix(_f64(key, p + 32) ^ RS[2], _f64(key, p + 40) ^ s2)
                p += 48
                i -= 48
                if i < 48:
                    break
            seed ^= s1 ^ s2
        if i > 16:
            seed = _mix(_f64(key, p) ^ RS[2], _f64(key, p + 8) ^ seed ^ RS[1])
            if i > 32:
                seed = _mix(_f64(key, p + 16) ^ RS[2], _f64(key, p + 24) ^ seed)
        a, b = _f64(key, n - 16), _f64(key, n - 8)
    a ^= RS[1]
    b ^= seed
    a, b = _mum(a, b)
    return f"{_mix(a ^ RS[0] ^ n, b ^ RS[final_secret]):016x}"
