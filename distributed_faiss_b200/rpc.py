"""Control-plane RPC: call methods of a remote IndexServer by name.

Same programming model as the reference's `rpc.py` (distributed_faiss/rpc.py:104-138): a
`Client` proxy whose attribute access turns into `(method_name, args)` sent to the server,
the reply is `(error_text_or_None, return_value)` and a non-None error is re-raised as
`ServerException` (rpc.py:126-131).  The wire format is this package's own: every message is
an 8-byte big-endian length followed by a pickle (protocol 4) -- one `sendall`, exact-size
reads -- instead of the reference's unframed pickle stream.

On an 8xB200 box this channel carries only the CONTROL plane (create / add / train / state /
save) and, for API compatibility, socket-mode search.  The timed search data plane is NCCL
(distributed_faiss_b200/spmd.py), which replaces the per-shard pickled query/result traffic
described in SURVEY.md section 5.
"""
import pickle
import socket
import struct

DEFAULT_PORT = 12032
_HDR = struct.Struct(">Q")


class ClientExit(Exception):
    """The peer closed the connection."""


class ServerException(Exception):
    """An exception raised inside the server while executing a call."""


def send_msg(sock: socket.socket, obj) -> None:
    payload = pickle.dumps(obj, protocol=4)
    sock.sendall(_HDR.pack(len(payload)))
    sock.sendall(payload)


def _recv_exact(sock: socket.socket, n: int) -> bytearray:
    buf = bytearray(n)
    view = memoryview(buf)
    got = 0
    while got < n:
        r = sock.recv_into(view[got:], min(n - got, 1 << 26))
        if r == 0:
            raise ClientExit("connection closed")
        got += r
    return buf


def recv_msg(sock: socket.socket):
    (n,) = _HDR.unpack(_recv_exact(sock, _HDR.size))
    return pickle.loads(_recv_exact(sock, n))


class Client:
    """Proxy of one IndexServer: `client.search(index_id, q, k, False)` runs there."""

    def __init__(self, id, HOST, port=DEFAULT_PORT, v6=False):
        self.id = id
        family = socket.AF_INET6 if v6 else socket.AF_INET
        self.sock = socket.socket(family, socket.SOCK_STREAM)
        self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self.sock.connect((HOST, port))

    def generic_fun(self, fname, args):
        send_msg(self.sock, (fname, args))
        return self.get_result()

    def get_result(self):
        st, ret = recv_msg(self.sock)
        if st is not None:
            raise ServerException(st)
        return ret

    def close(self):
        try:
            self.sock.shutdown(socket.SHUT_RDWR)
        except OSError:
            pass
        self.sock.close()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *args: self.generic_fun(name, args)
