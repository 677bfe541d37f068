"""Generates tests/golden/dxbc_literals.json from the reference's shipped shader blobs.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_dxbc_literals.py
For each hot-path blob (Bin/CSAdvect.cso, CSProject3D.cso, CSProject2D.cso) it walks the DXBC
container, takes the SHEX chunk and records every aligned 32-bit word of it that is the fp32 bit
pattern of a literal the restatement uses.  The test (tests/test_oracle.py::test_constants_match_dxbc)
then requires the oracle's constants to be exactly these bit patterns.
"""
import hashlib
import json
import os
import struct

REF = "/root/reference/Bin"
WANT = {
    "CSAdvect.cso": ["0x3f000000", "0xbdcccccd", "0xc3480000", "0x43480000", "0xc0800000", "0x3b800000", "0x3a800000",
                     "0x3fb8aa3b", "0x43400000", "0x42400000", "0x3c960aae", "0x41000000", "0x41800000", "0x42200000",
                     "0x3e4ccccd", "0x3f800000"],
    "CSProject3D.cso": ["0x3f000000", "0x3e2aaaab", "0x3a83126f", "0x3f855556", "0x40000000", "0xbf800000",
                        "0x3f7851ec", "0x42055556", "0x3f800000"],
    "CSProject2D.cso": ["0x3f000000", "0x3e800000", "0x3a83126f", "0x40000000", "0xbf800000", "0x3f7851ec",
                        "0x42055556", "0x3f800000"],
}


def shex_words(blob: bytes):
    assert blob[:4] == b"DXBC"
    n_chunks = struct.unpack_from("<I", blob, 28)[0]
    offsets = struct.unpack_from("<%dI" % n_chunks, blob, 32)
    for off in offsets:
        tag = blob[off:off + 4]
        size = struct.unpack_from("<I", blob, off + 4)[0]
        if tag in (b"SHEX", b"SHDR"):
            data = blob[off + 8:off + 8 + size]
            return struct.unpack("<%dI" % (len(data) // 4), data)
    raise RuntimeError("no SHEX chunk")


def main():
    out = {}
    for name, want in WANT.items():
        blob = open(os.path.join(REF, name), "rb").read()
        words = set(shex_words(blob))
        found = {w: (int(w, 16) in words) for w in want}
        out[name] = {"sha256": hashlib.sha256(blob).hexdigest(), "shex_tokens": len(shex_words(blob)),
                     "literals_found": found}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dxbc_literals.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
