function ampliflat(x, atype, options)
%AMPLIFLAT  Flat-gain optical amplifier with ASE noise (drop-in front-end, device side).
%   AMPLIFLAT(X,ATYPE,OPTIONS) keeps the contract of the toolbox's ampliflat.m: the field in
%   GSTATE.FIELDX / FIELDY is multiplied by sqrt(gain), gain = 10^(X/10) for ATYPE 'gain' and
%   X/avg_power(midch,'abs') for 'fixpower' (separate channels only), and, with OPTIONS.f [dB],
%   complex Gaussian noise of the amplifier's ASE is added on both polarizations (OPTIONS.onepol =
%   'asex' | 'asey' restricts it to one; OPTIONS.noise injects the noise samples, Nfft x 2*nfc).
%   The arithmetic runs through the MEX gateway on the field the previous in-line device left in HBM,
%   so a loop  fiber(...); ampliflat(...);  uploads the field once.  Without OPTIONS.noise the ASE
%   comes from the device generator, seeded from the interpreter's RAND stream (one draw per call),
%   so that rand('state',k) makes a run repeatable and every call adds independent noise.

global CONSTANTS GSTATE PMXOPT

ncol = size(GSTATE.FIELDX, 2);
switch lower(atype)
    case 'gain'
        gain = 10^(x * 0.1);
    case 'fixpower'
        if ncol ~= GSTATE.NCH
            error(['''fixpower'' works', ' only for channels separated']);
        end
        gain = x / avg_power(ceil(ncol / 2), 'abs');   % (the toolbox's own avg_power, on the write-through copy)
    otherwise
        error('wrong string atype');
end
sigma = zeros(1, ncol);
asepol = 3;
noise = [];
if nargin > 2 && isstruct(options)
    if isfield(options, 'f') && ~isinf(options.f)
        flin = 10^(options.f * 0.1);
        if ncol == 1
            maxl = max(GSTATE.LAMBDA);
            minl = min(GSTATE.LAMBDA);
            lam = 2 * maxl * minl / (maxl + minl);
        else
            lam = GSTATE.LAMBDA;
        end
        sigma = sqrt(flin / 4 * CONSTANTS.HPLANCK * CONSTANTS.CLIGHT ./ lam * (gain - 1) * GSTATE.NT * ...
            GSTATE.SYMBOLRATE * 1e21);
    end
    if isfield(options, 'onepol')
        if strcmp(options.onepol, 'asex')
            asepol = 1;
        elseif strcmp(options.onepol, 'asey')
            asepol = 2;
        else
            error('ONEPOL, if exists, must be ''asex'' or ''asey''');
        end
    end
    if isfield(options, 'noise')
        noise = options.noise;
    end
end
ase = any(sigma ~= 0);
if ase && isempty(noise)
    noise = floor(rand(1) * 2^52);               % seed of the device generator for this call
end
prec = 0;
resident = 1;
if isstruct(PMXOPT)
    if isfield(PMXOPT, 'precision') && strcmp(PMXOPT.precision, 'f32')
        prec = 1;
    end
    if isfield(PMXOPT, 'resident')
        resident = double(PMXOPT.resident ~= 0);
    end
end
hady = ~isempty(GSTATE.FIELDY);
uy = GSTATE.FIELDY;
if ~hady
    uy = zeros(size(GSTATE.FIELDX));
end
[ux, uy] = ssfm_mex('ampliflat', GSTATE.FIELDX, uy, gain, sigma, noise, asepol, [0, prec, resident]);
GSTATE.FIELDX = ux;
if hady
    GSTATE.FIELDY = uy;
elseif ase && asepol ~= 1                        % the noise creates the second polarization
    GSTATE.FIELDY = uy;
    GSTATE.DELAY = [GSTATE.DELAY(1, :); zeros(1, GSTATE.NCH)];
end
