"""Turn a checkpoint written by Model.save_weights('*.npz') into the HDF5 file Keras' `model.load_weights('chkpt.hdf5')`
reads (/root/reference/train.py:99-100): one group per top-level layer with its `weight_names`, `layer_names` on the root
— the structure tf.keras' hdf5_format.save_weights_to_hdf5_group writes.  Needs h5py (NOT in the build image: this
script is untested there; the .npz holds exactly the names and arrays it copies).

    python tools/npz_to_keras_h5.py chkpt.npz chkpt.hdf5
"""
import sys

import numpy as np


def main(src, dst):
    import h5py
    with np.load(src, allow_pickle=False) as z, h5py.File(dst, "w") as f:
        layers = [str(n) for n in z["__layer_names__"]]
        f.attrs["layer_names"] = [n.encode("utf8") for n in layers]
        f.attrs["backend"] = b"tensorflow"
        f.attrs["keras_version"] = b"2.2.4-tf"
        for ln in layers:
            g = f.create_group(ln)
            names = [str(n) for n in z[f"__weight_names__/{ln}"]]
            g.attrs["weight_names"] = [n.encode("utf8") for n in names]
            for wn in names:
                g.create_dataset(wn, data=z[f"{ln}/{wn}"])
    print(f"{dst}: {len(layers)} layers")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
