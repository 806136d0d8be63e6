function varargout = inverse_pmd(brf, options)
%INVERSE_PMD  Inverse PMD matrix (drop-in front-end, device side).
%   INVERSE_PMD(BRF) applies the inverse of the PMD matrix (GVD included) of the link whose fibers returned the
%   structs in the cell BRF to GSTATE.FIELDX / FIELDY and clears GSTATE.DISP; [UINV,U] = INVERSE_PMD(BRF,OPTIONS)
%   also returns the [2,2,Nfft] matrices.  OPTIONS.gvd = 'no', OPTIONS.mat and OPTIONS.apply keep the meaning they
%   have in the toolbox's inverse_pmd.m (the field is transformed when OPTIONS.apply is absent or equal to 'n').
%   The inverse of a chain is the linear step of every fiber taken backwards with the plates in reverse order and all
%   phases negated: one single-step run of the split-step kernels per fiber through the MEX gateway; the matrices
%   are evaluated on the device as well, one thread per frequency.

global GSTATE

nfc = size(GSTATE.FIELDX, 2);
if nfc > 1, error('inverse_pmd can be used only with a unique field.'); end

isopt = exist('options', 'var');
gvd = ~(isopt && isfield(options, 'gvd') && strcmp(options.gvd, 'no'));
apply = ~isopt || ~isfield(options, 'apply') || strcmp(options.apply, 'n');
mat = [];
if isopt && isfield(options, 'mat')
    mat = options.mat;
end

nfiber = length(brf);
Nfft = length(GSTATE.FN);
ntr = zeros(1, nfiber);
lcorr = zeros(1, nfiber);
betat = zeros(Nfft, nfiber);
db1 = zeros(Nfft, nfiber);
plates = zeros(0, 3);
for n = 1:nfiber
    ntr(n) = length(brf{n}.theta);
    lcorr(n) = brf{n}.lcorr;
    betat(:, n) = brf{n}.betat(:);
    db1(:, n) = brf{n}.db1(:);
    plates = [plates; brf{n}.db0(:), brf{n}.theta(:), brf{n}.epsilon(:)];
end

if nargout >= 2
    [ux, uy, Uinv4, U4] = ssfm_mex('invpmd', GSTATE.FIELDX, GSTATE.FIELDY, plates, ntr, lcorr, betat, db1, mat, [gvd, apply]);
    varargout{1} = reshape(Uinv4, 2, 2, Nfft);
    varargout{2} = reshape(U4, 2, 2, Nfft);
elseif nargout == 1
    [ux, uy, Uinv4] = ssfm_mex('invpmd', GSTATE.FIELDX, GSTATE.FIELDY, plates, ntr, lcorr, betat, db1, mat, [gvd, apply]);
    varargout{1} = reshape(Uinv4, 2, 2, Nfft);
else
    [ux, uy] = ssfm_mex('invpmd', GSTATE.FIELDX, GSTATE.FIELDY, plates, ntr, lcorr, betat, db1, mat, [gvd, apply]);
end
if apply
    GSTATE.FIELDX = ux;
    GSTATE.FIELDY = uy;
    GSTATE.DISP = zeros(2, GSTATE.NCH);
end
