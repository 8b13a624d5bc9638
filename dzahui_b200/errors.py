"""`Error` of the reference (error.rs:32-55), as exceptions; only the variants the hot path raises."""


class Error(Exception):
    """Base of every error the library reports (reference: `enum Error`)."""
    code = -1


class WrongDims(Error):        # Error::WrongDims
    code = 1


class Integration(Error):      # Error::Integration(String)
    code = 2


class PieceWiseDims(Error):    # Error::PieceWiseDims
    code = 3


class MeshParse(Error):        # Error::MeshParse(String)
    code = 4


class Io(Error):               # Error::Io
    code = 5


class CudaError(Error):        # no reference equivalent: CUDA failure / no device (there is no CPU fallback)
    code = 6


class Invalid(Error):          # no reference equivalent: bad handle or option
    code = 7


_BY_CODE = {c.code: c for c in (WrongDims, Integration, PieceWiseDims, MeshParse, Io, CudaError, Invalid)}


def check(rc, lib):
    if rc == 0:
        return
    msg = lib.dzb_last_error()
    msg = msg.decode("utf-8", "replace") if msg else ""
    raise _BY_CODE.get(rc, Error)(msg)
