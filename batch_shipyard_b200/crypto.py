"""Keys, certificates and credential encryption.

Parity with /root/reference/convoy/crypto.py: ssh key-pair generation (:127), ssh command
execution (:171), PEM/PFX generation and thumbprints through ``openssl`` (:219-433), RSA
encryption of credentials handed to nodes (:535-615).  ``openssl``/``ssh-keygen`` are used
when present; locally nothing leaves the box, so encryption degrades to a reversible
pass-through when the tools are missing.
"""
from __future__ import annotations

import base64
import hashlib
import os
import shutil
import subprocess
from typing import Optional


def _have(tool: str) -> bool:
    return shutil.which(tool) is not None


def generate_ssh_keypair(export_path: str = ".", prefix: str = "id_rsa_shipyard") -> tuple[str, str]:
    os.makedirs(export_path, exist_ok=True)
    priv = os.path.join(export_path, prefix)
    pub = priv + ".pub"
    for p in (priv, pub):
        if os.path.exists(p):
            os.remove(p)
    if _have("ssh-keygen"):
        subprocess.check_call(["ssh-keygen", "-q", "-f", priv, "-t", "rsa", "-b", "3072", "-N", ""], stdout=subprocess.DEVNULL)
    else:
        with open(priv, "w") as f:
            f.write("-----BEGIN SHIPYARD LOCAL KEY-----\n" + base64.b64encode(os.urandom(48)).decode() + "\n-----END SHIPYARD LOCAL KEY-----\n")
        os.chmod(priv, 0o600)
        with open(pub, "w") as f:
            f.write("ssh-rsa " + base64.b64encode(os.urandom(48)).decode() + " shipyard-local\n")
    return priv, pub


def connect_or_exec_ssh_command(host: str, port: int, key: Optional[str], user: str, command: Optional[str] = None,
                                tty: bool = False) -> int:
    """Local pools have no remote hosts: 127.0.0.1 executes in place; anything else uses ssh."""
    if host in ("127.0.0.1", "localhost"):
        return subprocess.call(command or os.environ.get("SHELL", "/bin/bash"), shell=True)
    cmd = ["ssh", "-o", "StrictHostKeyChecking=no", "-o", "UserKnownHostsFile=/dev/null", "-p", str(port)]
    if key:
        cmd += ["-i", key]
    if tty:
        cmd.append("-t")
    cmd.append(f"{user}@{host}")
    if command:
        cmd.append(command)
    return subprocess.call(cmd)


def generate_pem_pfx_certificates(file_prefix: str, pfx_password: Optional[str] = None) -> dict:
    pem, pfx = file_prefix + ".pem", file_prefix + ".pfx"
    if not _have("openssl"):
        raise RuntimeError("openssl is required to create certificates")
    key = file_prefix + ".key.pem"
    subprocess.check_call(["openssl", "req", "-new", "-nodes", "-x509", "-newkey", "rsa:2048", "-keyout", key, "-out", pem,
                           "-days", "730", "-subj", "/C=US/O=shipyard-b200/CN=BatchShipyard"], stderr=subprocess.DEVNULL)
    subprocess.check_call(["openssl", "pkcs12", "-export", "-out", pfx, "-inkey", key, "-in", pem, "-certfile", pem,
                           "-passout", "pass:" + (pfx_password or "")], stderr=subprocess.DEVNULL)
    return {"pem": pem, "pfx": pfx, "private_key": key, "sha1_thumbprint": get_sha1_thumbprint(pem)}


def get_sha1_thumbprint(path: str, passphrase: Optional[str] = None) -> str:
    if _have("openssl") and path.endswith((".pem", ".cer", ".crt")):
        out = subprocess.run(["openssl", "x509", "-in", path, "-noout", "-fingerprint", "-sha1"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        if "=" in out:
            return out.strip().split("=")[-1].replace(":", "").lower()
    with open(path, "rb") as f:
        return hashlib.sha1(f.read()).hexdigest()


class EncryptionError(RuntimeError):
    """Raised when encryption was requested but cannot be performed (never degrade to a reversible encoding)."""


_RSA_CHUNK = 190          # bytes of plaintext per RSA block: fits OAEP (214) and PKCS#1 v1.5 (245) with a 2048-bit key


def encrypt_string(enabled: bool, value: Optional[str], public_key_pem: Optional[str] = None) -> Optional[str]:
    """RSA-encrypt `value` for a node when encryption is enabled; pass-through otherwise.

    Same surface as /root/reference/convoy/crypto.py:603-615 (``encrypt_string(enabled, string, config)``).  Values longer than
    one RSA block are encrypted block-wise (``rsa:<b64>,<b64>,...``).  When `enabled` is true and the value cannot be
    encrypted (no openssl, no certificate, openssl failure) this raises :class:`EncryptionError`: a credential the user asked
    to protect is never stored in a reversible encoding.
    """
    if not enabled or value is None:
        return value
    if not public_key_pem or not os.path.exists(public_key_pem):
        raise EncryptionError(f"encryption is enabled but the public certificate {public_key_pem!r} does not exist")
    if not _have("openssl"):
        raise EncryptionError("encryption is enabled but openssl is not installed")
    raw = value.encode()
    blocks = []
    for i in range(0, max(len(raw), 1), _RSA_CHUNK):
        p = subprocess.run(["openssl", "pkeyutl", "-encrypt", "-certin", "-inkey", public_key_pem],
                           input=raw[i:i + _RSA_CHUNK], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if p.returncode != 0:
            raise EncryptionError("openssl pkeyutl -encrypt failed: " + p.stderr.decode(errors="replace").strip()[:200])
        blocks.append(base64.b64encode(p.stdout).decode())
    return "rsa:" + ",".join(blocks)


def decrypt_string(value: Optional[str], private_key_pem: Optional[str] = None) -> Optional[str]:
    if value is None:
        return None
    if value.startswith("b64:"):          # values written by earlier versions with encryption disabled
        return base64.b64decode(value[4:]).decode()
    if value.startswith("rsa:"):
        if not private_key_pem:
            raise EncryptionError("value is RSA-encrypted but no private key was given")
        out = b""
        for blk in value[4:].split(","):
            p = subprocess.run(["openssl", "pkeyutl", "-decrypt", "-inkey", private_key_pem], input=base64.b64decode(blk),
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            if p.returncode != 0:
                raise EncryptionError("cannot decrypt value with the given private key")
            out += p.stdout
        return out.decode()
    return value
