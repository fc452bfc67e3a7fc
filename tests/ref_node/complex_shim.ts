// Stand-in for the missing src/math/complex.ts (imported by src/math/vector.ts:17, used only by to_complex()).
export class Complex {
	re: number;
	im: number;
	constructor(re: number, im: number) { this.re = re; this.im = im; }
}
