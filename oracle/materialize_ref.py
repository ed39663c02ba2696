"""TEST INFRASTRUCTURE: materialise the reference's own Python sources of the hot path into oracle/_ref/ (git-ignored, NOT
gpurun-ignored), so the GPU box can run the UNMODIFIED reference modules as the timed CPU baseline.  The reference is 100 %
Python: there is nothing to compile -- "building" oracle/_ref is a file copy of stllm/models/*.py and
stllm/common/*.py (the files oracle/ref_shim.py loads per file).  Run by __graft_entry__.build() wherever /root/reference
exists; a no-op elsewhere.  No reference source is committed to this repository."""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
SUBDIRS = ("stllm/models", "stllm/common")


def materialize(src=SRC, dst=os.path.join(HERE, "_ref")):
    if not os.path.isfile(os.path.join(src, "stllm/models/st_llm.py")):
        return False
    n = 0
    for sub in SUBDIRS:
        os.makedirs(os.path.join(dst, sub), exist_ok=True)
        for fn in sorted(os.listdir(os.path.join(src, sub))):
            if fn.endswith(".py"):
                shutil.copyfile(os.path.join(src, sub, fn), os.path.join(dst, sub, fn))
                n += 1
    with open(os.path.join(dst, "README"), "w") as f:
        f.write("Copied from the reference tree by oracle/materialize_ref.py (build artefact, git-ignored).\n")
    return n


if __name__ == "__main__":
    print("materialised", materialize(), "files into oracle/_ref")
