import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from oracle.oracle import OracleChunk

def run(backend, nx, ny, inner, cap, solver='ppcg'):
    s = classic_settings(nx, ny=ny, steps=1, solver=solver, ppcginnersteps=inner, maxiters=cap)
    chunk, geom = tl.initialiseapp(s, backend=backend)
    recs, final = tl.diffuse(chunk, s, geom)
    return recs[0], chunk

for (nx, ny, inner) in ((65, 70, 7), (96, 160, 4), (128,128,10)):
    print('case', nx, ny, inner)
    for cap in (30, 31, 32, 33, 35):
        rd, cd = run(DeviceChunk, nx, ny, inner, cap)
        ro, co = run(OracleChunk, nx, ny, inner, cap)
        du = np.abs(cd.get_field('u') - co.get_field('u')).max() / np.abs(co.get_field('u')).max()
        scale = np.abs(co.get_field('u')).max()
        dp = np.abs(cd.get_field('p') - co.get_field('p')).max() / scale
        dsd = np.abs(cd.get_field('sd') - co.get_field('sd')).max() / scale
        dr = np.abs(cd.get_field('r') - co.get_field('r')).max() / scale
        print(cap, 'dev', rd['iters'], rd['error'], 'ora', ro['iters'], ro['error'], 'du %.2e dp %.2e dsd %.2e dr %.2e' % (du, dp, dsd, dr),
              'eig', rd.get('eigmin'), ro.get('eigmin'), rd.get('eigmax'), ro.get('eigmax'))
