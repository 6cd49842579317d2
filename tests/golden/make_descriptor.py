"""Extracts the serialized FileDescriptorProto of isg_ai.proto from the reference's generated module (UNet/isg_ai_pb2.py:19-22) into
tests/golden/isg_ai_descriptor.bin.  The module itself cannot be imported under protobuf >= 4 ("Descriptors cannot be created
directly"), but its descriptor bytes load through descriptor_pool.AddSerializedFile -- which is how tests/test_reader_cpu.py pins
unetb200.imagereader.encode_pair / decode_pair against the reference's OWN message definition.
Run in the build container (needs /root/reference):  python tests/golden/make_descriptor.py"""
import ast
import os
import re

SRC = "/root/reference/UNet/isg_ai_pb2.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "isg_ai_descriptor.bin")

if __name__ == "__main__":
    text = open(SRC).read()
    m = re.search(r"serialized_pb=_b\(('(?:[^'\\]|\\.)*')\)", text)
    raw = ast.literal_eval(m.group(1)).encode("latin1")
    from google.protobuf import descriptor_pb2
    fdp = descriptor_pb2.FileDescriptorProto.FromString(raw)          # sanity: it parses, and it is the message we think it is
    assert fdp.name == "isg_ai.proto" and fdp.message_type[0].name == "ImageMaskPair" and len(fdp.message_type[0].field) == 8
    open(OUT, "wb").write(raw)
    print(f"wrote {len(raw)} bytes -> {OUT}")
