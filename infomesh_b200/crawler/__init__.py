"""Web crawler (CPU plane): fetch -> parse -> dedup (URL / SHA-256 / SimHash) -> schedule links."""
from __future__ import annotations

import ssl
from pathlib import Path

MAX_RESPONSE_BYTES: int = 10 * 1024 * 1024   # shared body-size cap for every fetch path
_CA_BUNDLES = ("/etc/ssl/certs/ca-certificates.crt", "/etc/pki/tls/certs/ca-bundle.crt", "/etc/ssl/cert.pem",
               "/etc/pki/ca-trust/extracted/pem/tls-ca-bundle.pem")


def create_ssl_context() -> ssl.SSLContext | bool:
    """TLS context from the first system CA bundle found; ``True`` lets httpx use certifi."""
    for p in _CA_BUNDLES:
        if Path(p).is_file():
            return ssl.create_default_context(cafile=p)
    return True
