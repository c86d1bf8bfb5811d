"""TEST/BUILD INFRASTRUCTURE.  Regenerates hypercraft_b200/data/registry.bin -- the flat `hcm_registry` blob
(include/hcm.h) -- by running the reference's own registry (oracle/_ref/libhcref.so, built from
/root/reference/block + /root/reference/texture).  The blob is DATA produced by the reference, committed so that
the product can run where /root/reference is absent; INTEGRATION.md shows the exporter a HyperCraft build would link."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hcoracle  # noqa: E402

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hypercraft_b200", "data", "registry.bin")
    blob = hcoracle.RefWorld().registry_blob()
    blob.tofile(out)
    print(f"wrote {os.path.normpath(out)}: {blob.size} bytes")
