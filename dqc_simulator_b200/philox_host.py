"""Host-side Philox4x32-10 draw with the engine's keying (seed; counter = global shot id, draw
index), for the few places where the HOST must reproduce a device-side random number: the sharded
state-vector measurement, where every rank draws the same outcome after the probability all-reduce.
Pure Python integers; same published algorithm as csrc/dqe_device.cuh (Salmon et al., SC'11)."""

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c = [int(x) & _MASK for x in ctr]
    k0, k1 = int(key[0]) & _MASK, int(key[1]) & _MASK
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c[3] ^ k1) & _MASK, p0 & _MASK]
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c


def _u53(hi, lo):
    return ((hi >> 5) * 67108864.0 + (lo >> 6)) * (1.0 / 9007199254740992.0)


def shot_uniforms(seed, shot, draw):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10([shot & _MASK, (shot >> 32) & _MASK, draw & _MASK, (draw >> 32) & _MASK],
                      [seed & _MASK, seed >> 32])
    return _u53(r[0], r[1]), _u53(r[2], r[3])
