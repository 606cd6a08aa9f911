"""Codec base: the attributes SDP negotiation reads (mirror of /root/reference/Core/Codecs/GenCodec.py:1-13)."""


class GenCodec:
    srate: int = 8000   # sample rate of the decoded audio
    crate: int = 8000   # RTP clock rate
    ptype: int          # RTP payload type
    ename: str          # encoding name in a=rtpmap

    def __init__(self):
        if getattr(self, "ptype", None) is None or getattr(self, "ename", None) is None:
            raise AssertionError("codec classes must define ptype and ename")

    @classmethod
    def rtpmap(cls) -> str:
        if not all(hasattr(cls, a) for a in ("ptype", "ename")):
            raise AssertionError("codec classes must define ptype and ename")
        return f"rtpmap:{cls.ptype} {cls.ename}/{cls.crate}"
