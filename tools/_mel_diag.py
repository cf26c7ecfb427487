import sys, torch
sys.path.insert(0,'/root/repo')
from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import Engine
from oracle import port
dev=torch.device('cuda',0)
sd=syn.synthetic_state_dict(0)
W=port.Weights(sd)
eng=Engine(dev,'fp32'); eng.load_state_dict({k:v for k,v in sd.items() if k.startswith('spectrogram.')})
for name,wave in (('noise',syn.audio_noise(2,0)),('tones',syn.audio_tones(2,0))):
    ref=port.logmel(wave,W.window,W.fb)
    ref64=port.logmel(wave,W.window,W.fb,dtype=torch.float64).float()
    eng.set_flags(mel='simt'); a=eng.logmel(wave.to(dev)).cpu()
    eng.set_flags(mel='tc'); b=eng.logmel(wave.to(dev)).cpu()
    for lab,x in (('simt',a),('tc',b)):
        d=(x-ref64).abs()
        top=torch.topk(d.flatten(),6)
        locs=[(int(i)//(188*384), (int(i)//384)%188, int(i)%384) for i in top.indices]
        print(name,lab,'max abs vs f64',float(d.max()),'mean',float(d.mean()),'top',[(l,round(float(v),5),round(float(ref64.flatten()[i]),3)) for l,v,i in zip(locs,top.values,top.indices)])
    d=(a-b).abs(); print(name,'simt vs tc max',float(d.max()),'mean',float(d.mean()))
