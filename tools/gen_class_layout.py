#!/usr/bin/env python3
"""Class layouts and vtables of the public headers, as g++ lays them out (-fdump-lang-class), in a normal form that can
be compared between two header trees.

    python tools/gen_class_layout.py                 # freeze the REFERENCE's layout into tests/golden/ref_class_layout_v1.txt
    python tools/gen_class_layout.py --print ROOT    # print the layout of the header tree under ROOT

The fixture is what tests/test_cpp_api.py::test_vtable_layout_matches_reference holds this build's headers
(image_compression_b200/cpp) against on a machine without /root/reference: member offsets and sizes of
CompressedImage / Compressor / DxtcCompressor / EtcCompressor / PvrtcCompressor and the order of the virtuals in each
vtable.  An object file compiled against the reference's headers dispatches correctly into this build's library iff
these agree (public/compressor.h:52-137 of the reference declares the order).
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
FIXTURE = os.path.join(ROOT, "tests", "golden", "ref_class_layout_v1.txt")
HEADERS = ["compressed_image", "compressor", "dxtc_compressor", "etc_compressor", "pvrtc_compressor", "dxtc_to_etc_transcoder"]


def layout_of(include_root, cxx="g++"):
    include_root = os.path.abspath(include_root)
    """Normalised text: every 'Vtable for' / 'Class' paragraph of namespace image_codec_compression, addresses removed."""
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for h in HEADERS:
            src = os.path.join(tmp, h + ".cc")
            with open(src, "w") as f:
                f.write('#include "image_compression/public/%s.h"\n' % h)
            dump = os.path.join(tmp, h + ".class")
            subprocess.run([cxx, "-std=c++14", "-DIS_LITTLE_ENDIAN", "-I", include_root, "-fdump-lang-class=" + dump, "-c", src,
                            "-o", os.devnull], check=True, cwd=tmp)
            text = open(dump).read() if os.path.exists(dump) else ""
            for para in text.split("\n\n"):
                head = para.strip().split("\n")[0]
                if "image_codec_compression" not in head or not (head.startswith("Vtable for") or head.startswith("Class")):
                    continue
                para = re.sub(r"\(0x0x[0-9a-f]+\)", "", para.strip())
                if para not in out:
                    out.append(para)
    return "\n\n".join(out) + "\n"


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--print":
        sys.stdout.write(layout_of(sys.argv[2]))
        return
    if not os.path.isdir(REFERENCE):
        raise SystemExit("%s is not mounted; the fixture can only be regenerated where the reference is" % REFERENCE)
    text = layout_of(REFERENCE)
    with open(FIXTURE, "w") as f:
        f.write(text)
    print("wrote %s (%d paragraphs)" % (FIXTURE, text.count("\n\n") + 1))


if __name__ == "__main__":
    main()
