"""SSRF guard for every outbound fetch (crawler, MCP ``fetch_page`` / ``crawl_url``, webhooks).

Checks and block lists follow reference infomesh/security.py:25-146: http(s) only, URL length cap, cloud-metadata and
``*.local / *.internal / *.intranet / localhost`` hostnames, private / reserved / multicast IPv4+IPv6 literals, and
(optionally, and always after redirects) the DNS-resolved addresses.
"""
from __future__ import annotations

import ipaddress
import re
import socket
from urllib.parse import urlparse

MAX_URL_LENGTH = 2048
_SCHEMES = frozenset({"http", "https"})
_BLOCKED_NETS = tuple(ipaddress.ip_network(n) for n in (
    "0.0.0.0/8", "10.0.0.0/8", "127.0.0.0/8", "169.254.0.0/16", "172.16.0.0/12", "192.0.0.0/24", "192.0.2.0/24",
    "192.168.0.0/16", "198.18.0.0/15", "198.51.100.0/24", "203.0.113.0/24", "224.0.0.0/4", "240.0.0.0/4",
    "255.255.255.255/32", "::1/128", "fc00::/7", "fe80::/10", "ff00::/8"))
_BLOCKED_NAME = re.compile(r"^(localhost|.*\.local|.*\.internal|.*\.intranet|metadata\.google\.internal)$", re.I)
_BLOCKED_HOSTS = frozenset({"metadata.google.internal", "169.254.169.254", "[fd00:ec2::254]"})


class SSRFError(Exception):
    """The URL targets something a crawler must never touch."""


def _is_blocked_ip(ip: ipaddress.IPv4Address | ipaddress.IPv6Address) -> bool:
    if isinstance(ip, ipaddress.IPv6Address) and ip.ipv4_mapped is not None:
        ip = ip.ipv4_mapped
    return any(ip.version == net.version and ip in net for net in _BLOCKED_NETS)


def _check_resolved(hostname: str) -> None:
    try:
        infos = socket.getaddrinfo(hostname, None, socket.AF_UNSPEC, socket.SOCK_STREAM)
    except socket.gaierror as exc:
        raise SSRFError(f"DNS resolution failed for '{hostname}': {exc}") from exc
    for *_, sockaddr in infos:
        try:
            ip = ipaddress.ip_address(sockaddr[0])
        except ValueError:
            continue
        if _is_blocked_ip(ip):
            raise SSRFError(f"Hostname '{hostname}' resolves to private IP {ip}")


def validate_url(url: str, *, resolve_dns: bool = False) -> str:
    if not url or not isinstance(url, str):
        raise SSRFError("Empty or invalid URL")
    if len(url) > MAX_URL_LENGTH:
        raise SSRFError(f"URL exceeds maximum length of {MAX_URL_LENGTH}")
    try:
        parsed = urlparse(url)
        host = parsed.hostname
    except ValueError as exc:
        raise SSRFError(f"Malformed URL: {exc}") from exc
    if parsed.scheme not in _SCHEMES:
        raise SSRFError(f"Scheme '{parsed.scheme}' not allowed; must be one of {sorted(_SCHEMES)}")
    if not host:
        raise SSRFError("URL has no hostname")
    if host in _BLOCKED_HOSTS:
        raise SSRFError(f"Hostname '{host}' is blocked (metadata endpoint)")
    if _BLOCKED_NAME.match(host):
        raise SSRFError(f"Hostname '{host}' matches blocked pattern")
    try:
        literal = ipaddress.ip_address(host)
    except ValueError:
        literal = None
    if literal is not None and _is_blocked_ip(literal):
        raise SSRFError(f"IP address {literal} is in a private/reserved range")
    if resolve_dns and literal is None:
        _check_resolved(host)
    return url


def validate_url_post_redirect(final_url: str) -> str:
    """Re-validate the landing URL of a redirect chain, with DNS resolution (DNS-rebinding defence)."""
    return validate_url(final_url, resolve_dns=True)


def is_safe_url(url: str, *, resolve_dns: bool = False) -> bool:
    try:
        validate_url(url, resolve_dns=resolve_dns)
        return True
    except SSRFError:
        return False
