"""Image / map file formats on either side of the hot path (reference: src/main.jl:577-704).

  load_image(filename, ndims)   .mat (first N-D array, sorted variable names), .nii / .nii.gz (NIfTI-1, the
                                reference's scl_slope rule), always returned as a Fortran-ordered Float64
                                array so that the voxel index is the fastest one, as the C ABI expects
  save_mat(filename, dict)      MAT-file writer for the .t2dist.mat / .t2maps.mat / .t2parts.mat outputs
  save_nifti(filename, array)   minimal NIfTI-1 writer (used by the tests and for handing maps to viewers)

The reference writes MAT v7.3 (HDF5) through MAT.jl; this image has no HDF5 bindings, so the writer emits
MAT v5 (scipy.io.savemat) — same variable names, same array shapes, readable by MATLAB, MAT.jl and scipy.
PAR/REC/XML (Philips) inputs are not read here.
"""
import gzip
import os
import struct
import warnings

import numpy as np

ALLOWED_FILE_SUFFIXES = (".mat", ".nii", ".nii.gz", ".par", ".xml", ".rec")
ALLOWED_FILE_SUFFIXES_STRING = ", ".join(ALLOWED_FILE_SUFFIXES[:-1]) + ", and " + ALLOWED_FILE_SUFFIXES[-1]

# NIfTI-1 datatype codes -> numpy dtypes
_NIFTI_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4", 1024: "i8", 1280: "u8"}
_NIFTI_CODES = {np.dtype(v).str[1:]: k for k, v in _NIFTI_DTYPES.items()}


def maybe_get_suffix(filename):
    """Case-insensitive; the first allowed suffix that matches (src/main.jl:692-693).  `.nii.gz` files match
    `.nii.gz` because `.nii` does not end the name."""
    low = filename.lower()
    for ext in ALLOWED_FILE_SUFFIXES:
        if low.endswith(ext):
            return ext
    return None


def is_allowed_suffix(filename):
    return maybe_get_suffix(filename) is not None


def chop_allowed_suffix(filename):
    ext = maybe_get_suffix(filename)
    if ext is None:
        raise ValueError(f"Currently only {ALLOWED_FILE_SUFFIXES_STRING} file types are supported")
    return filename[: len(filename) - len(ext)]


def read_nifti(filename):
    """-> (raw array in file dtype, Fortran order; scl_slope; scl_inter).  NIfTI-1 single-file (.nii[.gz])."""
    opener = gzip.open if filename.lower().endswith(".gz") else open
    with opener(filename, "rb") as fh:
        buf = fh.read()
    if len(buf) < 348:
        raise ValueError(f"{filename}: not a NIfTI-1 file (shorter than its header)")
    endian = "<"
    if struct.unpack("<i", buf[:4])[0] != 348:
        if struct.unpack(">i", buf[:4])[0] != 348:
            raise ValueError(f"{filename}: not a NIfTI-1 file (sizeof_hdr != 348)")
        endian = ">"
    dim = struct.unpack(endian + "8h", buf[40:56])
    datatype, bitpix = struct.unpack(endian + "hh", buf[70:74])
    vox_offset, scl_slope, scl_inter = struct.unpack(endian + "3f", buf[108:120])
    magic = buf[344:348]
    if magic not in (b"n+1\0", b"ni1\0"):
        raise ValueError(f"{filename}: bad NIfTI-1 magic {magic!r}")
    if magic == b"ni1\0":
        raise ValueError(f"{filename}: header/image pairs (.hdr/.img) are not supported")
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{filename}: unsupported NIfTI datatype code {datatype}")
    nd = dim[0]
    if not 1 <= nd <= 7:
        raise ValueError(f"{filename}: bad dim[0] = {nd}")
    shape = tuple(int(d) for d in dim[1 : nd + 1])
    dt = np.dtype(endian + _NIFTI_DTYPES[datatype])
    count = int(np.prod(shape))
    off = int(vox_offset) if vox_offset >= 352 else 352
    raw = np.frombuffer(buf, dtype=dt, count=count, offset=off).reshape(shape, order="F")
    return raw, float(scl_slope), float(scl_inter)


def save_nifti(filename, array, scl_slope=1.0, scl_inter=0.0, pixdim=(1.0, 1.0, 1.0)):
    """Minimal NIfTI-1 single-file writer (`.nii`, gzip-compressed when the name ends in `.gz`)."""
    a = np.asarray(array)
    code = _NIFTI_CODES.get(a.dtype.str[1:])
    if code is None:
        a = a.astype(np.float64)
        code = 64
    if not 1 <= a.ndim <= 7:
        raise ValueError("NIfTI arrays have 1 to 7 dimensions")
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<hh", hdr, 70, code, a.dtype.itemsize * 8)
    pix = [1.0] + list(pixdim) + [1.0] * (7 - len(pixdim))
    struct.pack_into("<8f", hdr, 76, *pix[:8])
    struct.pack_into("<3f", hdr, 108, 352.0, scl_slope, scl_inter)
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + np.asfortranarray(a).astype(a.dtype.newbyteorder("<")).tobytes(order="F")
    opener = gzip.open if filename.lower().endswith(".gz") else open
    with opener(filename, "wb") as fh:
        fh.write(payload)


def ensure_ndims(filename, data, n):
    """src/main.jl:621-633: pad with trailing singleton dimensions, or select the first volume along extra ones."""
    d = data.ndim
    if d < n:
        return data.reshape(data.shape + (1,) * (n - d), order="F")
    if d == n:
        return data
    if any(s > 1 for s in data.shape[n:]):
        dims = ",".join(str(i) for i in range(n + 1, d + 1))
        warnings.warn(f"Input file {filename} has {d} dimensions, expected {n}; selecting the first {n}-D volume along "
                      f"{'dimension' if d - n == 1 else 'dimensions'} {dims}.")
    return data[(slice(None),) * n + (0,) * (d - n)]


def load_image(filename, ndims=4):
    """load_image(filename; ndims = 4)  (src/main.jl:577-619) -> Float64 array, Fortran order."""
    ext = maybe_get_suffix(filename)
    if ext == ".mat":
        try:
            from scipy.io import loadmat
            data = loadmat(filename)
        except NotImplementedError as e:  # MAT v7.3 is HDF5
            raise ValueError(f"{filename}: MAT v7.3 (HDF5) files need h5py, which this environment lacks; "
                             "re-save with -v7 or convert to NIfTI") from e
        keys = sorted(k for k, v in data.items() if not k.startswith("__") and isinstance(v, np.ndarray) and v.ndim == ndims
                      and v.dtype.kind in "fiub")
        if not keys:
            raise ValueError(f"No {ndims}-D array was found in the input file: {filename}")
        if len(keys) > 1:
            warnings.warn(f"Multiple possible images found in file: {filename}\nChoosing variable {keys[0]!r} out of the "
                          f"following options: {', '.join(repr(k) for k in keys)}")
        data = data[keys[0]]
    elif ext in (".nii", ".nii.gz"):
        raw, slope, inter = read_nifti(filename)
        if slope == 0:  # "if scl_slope == 0, data is not scaled and raw data should be returned"
            slope, inter = 1.0, 0.0
        # `nii.raw .* scl_slope .+ scl_inter` (src/main.jl:599) with Float32 header fields: Julia computes in
        # promote_type(eltype(raw), Float32) - Float32 for integer and Float32 volumes, Float64 for Float64 ones
        ct = np.float64 if raw.dtype == np.float64 else np.float32
        data = raw.astype(ct) * ct(slope) + ct(inter)
    elif ext in (".par", ".xml", ".rec"):
        raise ValueError(f"{filename}: Philips PAR/REC/XML files are not supported by this host mirror; convert to NIfTI")
    else:
        raise ValueError(f"Currently, only {ALLOWED_FILE_SUFFIXES_STRING} files are supported")
    data = ensure_ndims(filename, data, ndims)
    # "copyto!(Array{Float64, N}(undef, sz), data)": an owned, writable Float64 array (masks are applied in place)
    return np.require(np.asfortranarray(data, dtype=np.float64), requirements=["F", "W", "O"])


MAT5_MAX_BYTES = 2**31 - 2**20  # MAT v5 stores the byte count of a variable in 32 bits


def save_mat(filename, variables):
    """MAT.matwrite(savefile, dict): variable names and shapes as the reference writes them (MAT v5 container).
    A variable too large for MAT v5 (the reference writes v7.3 / HDF5, which this environment cannot) is written next
    to the file as `<filename>.<name>.npy` instead of failing after the whole volume has been computed; the .mat file
    then holds a string of that name pointing to it."""
    from scipy.io import savemat
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    out = {}
    for k, v in variables.items():
        a = v if np.isscalar(v) else np.asarray(v)
        if not np.isscalar(a) and a.nbytes > MAT5_MAX_BYTES:
            side = f"{filename}.{k}.npy"
            np.save(side, a)
            print(f"warning: {k} ({a.nbytes / 2**30:.1f} GiB) exceeds the MAT v5 limit; written to {side}")
            a = f"see {os.path.basename(side)}"
        out[k] = a
    savemat(filename, out, do_compression=False, oned_as="column")
