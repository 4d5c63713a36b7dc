"""Test helper: rewrite a baseline interleaved JPEG as a SEQUENTIAL file with one scan per component (SOF0, three
non-interleaved SOS segments) carrying the same quantised coefficients and the same Huffman tables.  Pillow cannot
write such files, libjpeg-turbo reads them, and the decoded pixels must equal those of the original file — a pin for
the multi-scan path of the decoders that does not involve progressive coding."""
import numpy as np

from oracle import jpeg_oracle

_ZZ = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49,
       56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def _segments(data):
    p, out = 2, []
    while True:
        assert data[p] == 0xFF
        m = data[p + 1]
        ln = (data[p + 2] << 8) | data[p + 3]
        out.append((m, data[p + 4:p + 2 + ln]))
        p += 2 + ln
        if m == 0xDA:
            return out


def _huff_codes(payload):
    """DHT payload -> {(class, id): {symbol: (code, length)}}"""
    tabs, o = {}, 0
    while o < len(payload):
        tc, th = payload[o] >> 4, payload[o] & 15
        bits = payload[o + 1:o + 17]
        n = sum(bits)
        vals = payload[o + 17:o + 17 + n]
        code, k, t = 0, 0, {}
        for ln in range(1, 17):
            for _ in range(bits[ln - 1]):
                t[vals[k]] = (code, ln)
                code += 1
                k += 1
            code <<= 1
        tabs[(tc, th)] = t
        o += 17 + n
    return tabs


class _Bits:
    def __init__(self):
        self.out, self.acc, self.n = bytearray(), 0, 0

    def put(self, v, ln):
        if not ln:
            return
        self.acc = (self.acc << ln) | (v & ((1 << ln) - 1))
        self.n += ln
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)
        return bytes(self.out)


def _mag(v):
    a = abs(int(v))
    s = a.bit_length()
    return s, (int(v) if v >= 0 else int(v) + (1 << s) - 1)


def one_scan_per_component(data: bytes) -> bytes:
    info = jpeg_oracle.info(data)
    assert info["components"] == 3 and info["restart_interval"] == 0
    segs = _segments(data)
    tabs = {}
    for m, pl in segs:
        if m == 0xC4:
            tabs.update(_huff_codes(pl))
    sos = [pl for m, pl in segs if m == 0xDA][0]
    sel = {sos[1 + 2 * i]: (sos[2 + 2 * i] >> 4, sos[2 + 2 * i] & 15) for i in range(3)}
    sof = [pl for m, pl in segs if m in (0xC0, 0xC1)][0]
    ids = [sof[6 + 3 * i] for i in range(3)]
    hmax = max(h for h, _ in info["sampling"])
    vmax = max(v for _, v in info["sampling"])
    out = bytearray(b"\xff\xd8")
    for m, pl in segs:
        if m != 0xDA:
            out += bytes([0xFF, m]) + (len(pl) + 2).to_bytes(2, "big") + bytes(pl)
    for ci in range(3):
        h, v = info["sampling"][ci]
        coef = jpeg_oracle.coefficients(data, ci)                      # [block rows][blocks per row][64], natural order
        bw = -(-(-(-info["width"] * h // hmax)) // 8)                   # the component's OWN block grid, not the MCU-padded one
        bh = -(-(-(-info["height"] * v // vmax)) // 8)
        td, ta = sel[ids[ci]]
        dc, ac = tabs[(0, td)], tabs[(1, ta)]
        bits, pred = _Bits(), 0
        for by in range(bh):
            for bx in range(bw):
                blk = coef[by, bx]
                s, val = _mag(int(blk[0]) - pred)
                pred = int(blk[0])
                bits.put(*dc[s])
                bits.put(val, s)
                run = 0
                for k in range(1, 64):
                    c = int(blk[_ZZ[k]])
                    if c == 0:
                        run += 1
                        continue
                    while run > 15:
                        bits.put(*ac[0xF0])
                        run -= 16
                    s, val = _mag(c)
                    bits.put(*ac[(run << 4) | s])
                    bits.put(val, s)
                    run = 0
                if run:
                    bits.put(*ac[0x00])
        out += b"\xff\xda" + (8).to_bytes(2, "big") + bytes([1, ids[ci], (td << 4) | ta, 0, 63, 0]) + bits.flush()
    return bytes(out + b"\xff\xd9")


# ---------------------------------------------------------------------------------------------------------------
# progressive files with ARBITRARY scan scripts (jcphuff.c's coding, without end-of-band runs longer than one block so
# that the Annex K tables of the source file suffice)
# ---------------------------------------------------------------------------------------------------------------
def _frame(data):
    info = jpeg_oracle.info(data)
    segs = _segments(data)
    tabs = {}
    for m, pl in segs:
        if m == 0xC4:
            tabs.update(_huff_codes(pl))
    sof = [pl for m, pl in segs if m in (0xC0, 0xC1)][0]
    ids = [sof[6 + 3 * i] for i in range(info["components"])]
    return info, segs, tabs, ids


def progressive_with_script(data: bytes, script) -> bytes:
    """Re-code a baseline 3-component file as SOF2 with the given scans: (components, Ss, Se, Ah, Al) each, components a
    tuple of indices (several only for DC scans).  The script has to send every bit of every coefficient eventually."""
    info, segs, tabs, ids = _frame(data)
    assert info["restart_interval"] == 0
    hmax = max(h for h, _ in info["sampling"])
    vmax = max(v for _, v in info["sampling"])
    W, H = info["width"], info["height"]
    mcux, mcuy = -(-W // (8 * hmax)), -(-H // (8 * vmax))
    coefs = [jpeg_oracle.coefficients(data, c) for c in range(3)]
    dc, ac = tabs[(0, 0)], tabs[(1, 0)]                      # the luma tables serve every scan
    out = bytearray(b"\xff\xd8")
    for m, pl in segs:
        if m in (0xC0, 0xC1):
            out += bytes([0xFF, 0xC2]) + (len(pl) + 2).to_bytes(2, "big") + bytes(pl)
        elif m != 0xDA:
            out += bytes([0xFF, m]) + (len(pl) + 2).to_bytes(2, "big") + bytes(pl)
    for comps, ss, se, ah, al in script:
        bits = _Bits()
        if ss == 0:                                          # DC scan: interleaved MCUs, or one component's own grid
            if len(comps) > 1:
                order = [(c, my * info["sampling"][c][1] + by, mx * info["sampling"][c][0] + bx)
                         for my in range(mcuy) for mx in range(mcux) for c in comps
                         for by in range(info["sampling"][c][1]) for bx in range(info["sampling"][c][0])]
            else:
                c = comps[0]
                h, v = info["sampling"][c]
                bw, bh = -(-(-(-W * h // hmax)) // 8), -(-(-(-H * v // vmax)) // 8)
                order = [(c, y, x) for y in range(bh) for x in range(bw)]
            pred = {c: 0 for c in comps}
            for c, y, x in order:
                d = int(coefs[c][y, x][0])
                if ah == 0:
                    t = d >> al                              # arithmetic shift, as jcphuff.c does for DC
                    s, val = _mag(t - pred[c])
                    pred[c] = t
                    bits.put(*dc[s])
                    bits.put(val, s)
                else:
                    bits.put((d >> al) & 1, 1)
        else:
            c = comps[0]
            h, v = info["sampling"][c]
            bw, bh = -(-(-(-W * h // hmax)) // 8), -(-(-(-H * v // vmax)) // 8)
            for y in range(bh):
                for x in range(bw):
                    blk = coefs[c][y, x]
                    if ah == 0:                              # first pass: magnitudes shifted towards zero
                        run = 0
                        for k in range(ss, se + 1):
                            cf = int(blk[_ZZ[k]])
                            t = abs(cf) >> al
                            if t == 0:
                                run += 1
                                continue
                            while run > 15:
                                bits.put(*ac[0xF0])
                                run -= 16
                            s = t.bit_length()
                            bits.put(*ac[(run << 4) | s])
                            bits.put(t if cf > 0 else (~t) & ((1 << s) - 1), s)
                            run = 0
                        if run:
                            bits.put(*ac[0x00])
                    else:                                    # refinement: correction bits ride behind the next symbol
                        absv = [abs(int(blk[_ZZ[k]])) >> al for k in range(64)]
                        eob = max([k for k in range(ss, se + 1) if absv[k] == 1], default=-1)
                        run, br = 0, []
                        for k in range(ss, se + 1):
                            t = absv[k]
                            if t == 0:
                                run += 1
                                continue
                            while run > 15 and k <= eob:
                                bits.put(*ac[0xF0])
                                run -= 16
                                for b in br:
                                    bits.put(b, 1)
                                br = []
                            if t > 1:
                                br.append(t & 1)
                                continue
                            bits.put(*ac[(run << 4) | 1])
                            bits.put(0 if int(blk[_ZZ[k]]) < 0 else 1, 1)
                            for b in br:
                                bits.put(b, 1)
                            br, run = [], 0
                        if run or br:
                            bits.put(*ac[0x00])
                            for b in br:
                                bits.put(b, 1)
        hdr = bytes([len(comps)]) + b"".join(bytes([ids[c], 0x00]) for c in comps) + bytes([ss, se, (ah << 4) | al])
        out += b"\xff\xda" + (len(hdr) + 2).to_bytes(2, "big") + hdr + bits.flush()
    return bytes(out + b"\xff\xd9")


# scan scripts that libjpeg's standard one does not cover
SCRIPTS = {
    "spectral selection only": [((0, 1, 2), 0, 0, 0, 0), ((0,), 1, 63, 0, 0), ((1,), 1, 63, 0, 0), ((2,), 1, 63, 0, 0)],
    "dc per component, deep approximation": [
        ((0,), 0, 0, 0, 2), ((1,), 0, 0, 0, 2), ((2,), 0, 0, 0, 2),
        ((0,), 1, 2, 0, 1), ((0,), 3, 63, 0, 2), ((1,), 1, 63, 0, 1), ((2,), 1, 63, 0, 1),
        ((0,), 3, 63, 2, 1), ((0,), 0, 0, 2, 1), ((1,), 0, 0, 2, 1), ((2,), 0, 0, 2, 1),
        ((0,), 1, 63, 1, 0), ((1,), 1, 63, 1, 0), ((2,), 1, 63, 1, 0),
        ((0,), 0, 0, 1, 0), ((1,), 0, 0, 1, 0), ((2,), 0, 0, 1, 0)],
    "fine bands refined one by one": [
        ((0, 1, 2), 0, 0, 0, 1),
        ((0,), 1, 5, 0, 1), ((0,), 6, 20, 0, 1), ((0,), 21, 63, 0, 1),
        ((1,), 1, 10, 0, 1), ((1,), 11, 63, 0, 1), ((2,), 1, 63, 0, 1),
        ((0, 1, 2), 0, 0, 1, 0),
        ((0,), 21, 63, 1, 0), ((0,), 1, 5, 1, 0), ((0,), 6, 20, 1, 0),
        ((1,), 1, 63, 1, 0), ((2,), 1, 30, 1, 0), ((2,), 31, 63, 1, 0)],
}


def random_script(rng, max_al=2):
    """A random VALID progression for 3 components: every band of every component gets a first pass at some Al and is
    then refined one bit at a time; bands, depths and the interleaving of the scans are drawn at random (per coefficient
    the order first pass -> refinements is kept, as jdphuff.c demands)."""
    chains = []                                             # each: list of scans that must stay in order
    al = int(rng.integers(0, max_al + 1))
    if rng.random() < 0.5:
        dc = [((0, 1, 2), 0, 0, 0, al)] + [((0, 1, 2), 0, 0, a + 1, a) for a in range(al - 1, -1, -1)]
        chains.append(dc)
    else:
        for c in range(3):
            a0 = int(rng.integers(0, max_al + 1))
            chains.append([((c,), 0, 0, 0, a0)] + [((c,), 0, 0, a + 1, a) for a in range(a0 - 1, -1, -1)])
    for c in range(3):
        cuts = sorted(set(int(x) for x in rng.integers(2, 63, int(rng.integers(0, 3)))))
        edges = [1] + cuts + [64]
        for lo, hi in zip(edges[:-1], edges[1:]):
            a0 = int(rng.integers(0, max_al + 1))
            chains.append([((c,), lo, hi - 1, 0, a0)] + [((c,), lo, hi - 1, a + 1, a) for a in range(a0 - 1, -1, -1)])
    script = []
    while chains:                                           # a random interleaving that keeps every chain's order
        k = int(rng.integers(0, len(chains)))
        script.append(chains[k].pop(0))
        if not chains[k]:
            chains.pop(k)
    # AC scans may not precede the first DC scan of their component (jdphuff.c checks coef_bits[0] >= 0): DC first passes lead
    head = [s for s in script if s[1] == 0 and s[3] == 0]
    rest = [s for s in script if not (s[1] == 0 and s[3] == 0)]
    return head + rest
