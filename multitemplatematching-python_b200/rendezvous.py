"""Hands the NCCL unique id of rank 0 to the other ranks of a torchrun-style launch -- without torch.

One process per GPU reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT from the environment
(the variables ``python -m torch.distributed.run`` sets).  Rank 0 listens on MASTER_ADDR:(MASTER_PORT + offset)
and sends the 128 id bytes to every rank that connects; the others retry until the listener is up.
``MTM_B200_COMM_PORT`` overrides the port.  Only the id crosses this socket: the data path is NCCL inside
libmtm_b200.so (``mtm_comm_init_rank``).
"""
import os
import socket
import time

from . import _native

_PORT_OFFSET = 117


def _endpoint(port_offset):
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    if "MTM_B200_COMM_PORT" in os.environ:
        return addr, int(os.environ["MTM_B200_COMM_PORT"]) + port_offset
    return addr, int(os.environ.get("MASTER_PORT", "29500")) + _PORT_OFFSET + port_offset


def exchange_id(rank, world, make_id, port_offset=0, timeout=120.0):
    """Returns the bytes ``make_id()`` produced on rank 0, on every rank."""
    if world == 1:
        return make_id()
    addr, port = _endpoint(port_offset)
    if rank == 0:
        payload = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("0.0.0.0" if addr not in ("127.0.0.1", "localhost") else "127.0.0.1", port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                conn, _peer = srv.accept()
                with conn:
                    conn.sendall(payload)
        finally:
            srv.close()
        return payload
    deadline = time.monotonic() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as s:
                chunks, need = [], _native.COMM_ID_BYTES
                while need:
                    part = s.recv(need)
                    if not part:
                        raise ConnectionError("rank 0 closed the id socket early")
                    chunks.append(part)
                    need -= len(part)
                return b"".join(chunks)
        except (ConnectionRefusedError, ConnectionError, socket.timeout, OSError):
            if time.monotonic() > deadline:
                raise
            time.sleep(0.05)


def comm_from_env(device=None, port_offset=0):
    """The communicator of a torchrun-style launch: one endpoint per process (``_native.Comm``)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if device is None:
        device = _native.local_device()
    uid = exchange_id(rank, world, _native.Comm.unique_id, port_offset) if world > 1 else None
    return _native.Comm.init_rank(device, world, rank, uid)
