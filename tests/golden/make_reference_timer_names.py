"""Generates tests/golden/reference_timer_names.txt: the timer names the reference registers (src/frame_handler_base.cpp:57-66), read from
/root/reference in the build container so that the GPU box (no reference tree) can still check hso_stage_name against them."""
import re
src = open("/root/reference/src/frame_handler_base.cpp").read()
names = re.findall(r'addTimer\("([a-z_]+)"\)', src)
open(__file__.rsplit("/", 1)[0] + "/reference_timer_names.txt", "w").write("\n".join(names) + "\n")
print(names)
