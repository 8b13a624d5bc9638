"""`Writer` of the reference (writer.rs:27-146): CSV snapshots of a solution vector, one file per call.
File I/O is host work; the formatting (Rust's `f64::to_string`) lives in the C-ABI library (csrc/csv_writer.cpp)."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import Error, Io, check


def format_f64(v):
    """`v.to_string()` as Rust prints an f64"""
    lib = _lib.load()
    buf = C.create_string_buffer(400)
    check(lib.dzb_format_f64(float(v), buf, len(buf)), lib)
    return buf.value.decode()


class Writer:
    def __init__(self, write_path, file_prefix, variable_names, erase_prev_dir=False):
        self._lib = _lib.load()
        self.write_path, self.file_prefix = str(write_path), str(file_prefix)
        self.variable_names = [str(v) for v in variable_names]
        rc = self._lib.dzb_csv_prepare_dir(self.write_path.encode(), 1 if erase_prev_dir else 0)
        if rc == Io.code:
            raise Io("Path for creating files and writing values not found")   # Error::NotFound (writer.rs:61-63)
        if rc:
            raise Error("dzb_csv_prepare_dir failed")

    def write(self, id, vals):
        vals = np.ascontiguousarray(np.asarray(vals, dtype=np.float64))
        names = (C.c_char_p * len(self.variable_names))(*[v.encode() for v in self.variable_names])
        rc = self._lib.dzb_csv_write(self.write_path.encode(), self.file_prefix.encode(), float(id), names, len(self.variable_names),
                                     vals.ctypes.data_as(_lib.c_double_p), vals.size)
        if rc == Io.code:
            raise Io("could not write the snapshot file")
        if rc:
            raise Error("dzb_csv_write failed")
        return f"{self.write_path}/{self.file_prefix}{format_f64(id)}.csv"
