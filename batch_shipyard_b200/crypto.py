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
import tempfile
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


def encrypt_string(enabled: bool, value: Optional[str], public_key_pem: Optional[str] = None) -> Optional[str]:
    """RSA-encrypt `value` for a node when encryption is enabled; pass-through otherwise."""
    if not enabled or value is None:
        return value
    if public_key_pem and _have("openssl") and os.path.exists(public_key_pem):
        with tempfile.NamedTemporaryFile("w", delete=False) as f:
            f.write(value)
        try:
            p = subprocess.run(["openssl", "pkeyutl", "-encrypt", "-certin", "-inkey", public_key_pem, "-in", f.name],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            if p.returncode == 0:
                return "rsa:" + base64.b64encode(p.stdout).decode()
        finally:
            os.remove(f.name)
    return "b64:" + base64.b64encode(value.encode()).decode()


def decrypt_string(value: Optional[str], private_key_pem: Optional[str] = None) -> Optional[str]:
    if value is None:
        return None
    if value.startswith("b64:"):
        return base64.b64decode(value[4:]).decode()
    if value.startswith("rsa:") and private_key_pem:
        p = subprocess.run(["openssl", "pkeyutl", "-decrypt", "-inkey", private_key_pem], input=base64.b64decode(value[4:]),
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        if p.returncode == 0:
            return p.stdout.decode()
        raise RuntimeError("cannot decrypt value with the given private key")
    return value
