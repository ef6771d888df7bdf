import re,glob,subprocess,sys,os
os.chdir(os.path.join(os.path.dirname(os.path.abspath(__file__)),'..','realtimeparticles_b200','lib'))
keys=sys.argv[1:]
for f in sorted(glob.glob('obj/*.ptxas.log')):
    txt=open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        name=subprocess.run(['c++filt',m.group(1)],capture_output=True,text=True).stdout.strip().split('(')[0]
        if not keys or any(k in name for k in keys):
            print(f"{name[:60]:60s} regs={m.group(5):>3s} stack={m.group(2)} spill={m.group(3)}/{m.group(4)}")
