"""Minimal FITS writer used when astropy is not installed (the reference writes
its optional `-f` outputs with astropy.io.fits, solex_util.py:204-206,
Solex_recon.py:80-82,137-152).  Host-side file output, outside the hot path.
If astropy is importable its Header / PrimaryHDU are used instead."""
from __future__ import annotations

import numpy as np

try:                                             # pragma: no cover - depends on the installation
    from astropy.io.fits import Header, PrimaryHDU   # noqa: F401
except Exception:

    class Header(dict):
        pass

    def _card(key, value, comment=''):
        if isinstance(value, bool):
            v = 'T' if value else 'F'
            body = '%-8s= %20s' % (key, v)
        elif isinstance(value, (int, np.integer)):
            body = '%-8s= %20d' % (key, int(value))
        elif isinstance(value, (float, np.floating)):
            body = '%-8s= %20.12G' % (key, float(value))
        else:
            body = "%-8s= '%-8s'" % (key, str(value).replace("'", "''"))
        return (body + (' / ' + comment if comment else ''))[:80].ljust(80)

    class PrimaryHDU:
        def __init__(self, data=None, header=None):
            self.data = None if data is None else np.asarray(data)
            self.header = Header(header or {})

        def writeto(self, path, overwrite=False):
            data = self.data
            cards = [_card('SIMPLE', True)]
            extra = []
            if data.dtype == np.uint16:              # stored as int16 with BZERO = 32768, as astropy does
                payload = (data.astype(np.int32) - 32768).astype('>i2')
                bitpix, extra = 16, [('BSCALE', 1), ('BZERO', 32768)]
            elif data.dtype.kind == 'f':
                payload, bitpix = data.astype('>f8'), -64
            else:
                payload, bitpix = data.astype('>i4'), 32
            cards += [_card('BITPIX', bitpix), _card('NAXIS', data.ndim)]
            for i, n in enumerate(reversed(data.shape)):
                cards.append(_card('NAXIS%d' % (i + 1), n))
            skip = {'SIMPLE', 'BITPIX', 'NAXIS', 'BZERO', 'BSCALE'} | {'NAXIS%d' % (i + 1) for i in range(data.ndim)}
            cards += [_card(k, v) for k, v in extra]
            cards += [_card(k, v) for k, v in self.header.items() if k not in skip]
            cards.append('END'.ljust(80))
            head = ''.join(cards).encode('ascii')
            head += b' ' * (-len(head) % 2880)
            body = payload.tobytes()
            body += b'\0' * (-len(body) % 2880)
            with open(path, 'wb') as f:
                f.write(head + body)
